"""Feature extraction + export (extract_feacture_v2.py:58-109).

`extract_features(net, data)` runs the host-framework backbone (any callable returning (output, fc) like the
reference's SymbolBlock, extract_feacture_v2.py:45-51,63) and L2-normalises fc on the GPU in one launch;
the writers reproduce the reference's text formats with ONE device->host copy per batch instead of one per
row."""
from __future__ import annotations

import numpy as np
import torch

from .norm import l2_normalize


def acc(output: torch.Tensor, label: torch.Tensor) -> float:
    """extract_feacture_v2.py:19-20: fraction of rows whose argmax equals the label."""
    return float((output.argmax(dim=1) == label.to(output.device).long()).float().mean())


@torch.no_grad()
def extract_features(net, data: torch.Tensor):
    """(output, fc) = net(data); returns (output, fc / ||fc||_2 per row)."""
    output, fc = net(data)
    return output, l2_normalize(fc.reshape(fc.shape[0], -1))


def format_feature_rows(fc_normalised: torch.Tensor) -> str:
    """Each row as '{},' per element (Python float repr of the fp32 value, trailing comma) + '\\n'
    (extract_feacture_v2.py:68-73)."""
    rows = fc_normalised.detach().to(torch.float32).cpu().numpy()
    return "".join("".join("{},".format(e) for e in row.tolist()) + "\n" for row in rows)


def format_label_rows(labels: torch.Tensor) -> str:
    """'{}\\n'.format(label[v].asscalar()) -> numpy.float32 str such as '12.0' (extract_feacture_v2.py:76-79)."""
    lab = torch.as_tensor(labels).detach().to(torch.float32).cpu().numpy().reshape(-1)
    return "".join("{}\n".format(np.float32(v)) for v in lab)


def write_feature_csv(path: str, fc_normalised: torch.Tensor, mode: str = "a+"):
    with open(path, mode, newline="") as f:
        f.write(format_feature_rows(fc_normalised))


def write_label_csv(path: str, labels: torch.Tensor, mode: str = "a+"):
    with open(path, mode, newline="") as f:
        f.write(format_label_rows(labels))
