"""Gallery scoring: the native recognition loop of
feature_extraction/c_version/Feature.hpp on the GPU.

  simd_dot                      Feature.hpp:273-293
  Compare_Face_From_DB          Feature.hpp:345-392 (argmax) / :295-343 (top-K)
  Compare_Face_Person / _DB     Feature.hpp:763-804
The gallery is stored SoA -- G (N, D) + gnorm (N,) (+ person ids) -- instead of the reference's AoS
FeatureImage/FeatureVec records (:105-158).
"""
from __future__ import annotations

import torch

from . import _lib as L
from .norm import row_norms


def simd_dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """float simd_dot(const float* x, const float* y, len): bit-identical fp32 result, shape ()."""
    L.require_cuda(x, y)
    x = x.reshape(1, -1).to(torch.float32).contiguous(); y = y.reshape(1, -1).to(torch.float32).contiguous()
    out = torch.empty((1, 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib.efm_simd_dot_rows(L.ptr(x), 1, x.shape[1], L.ptr(y), 1, L.ptr(out), L.stream()))
    return out.reshape(())


class Gallery:
    """Registered faces: rows `G`, stored norms `gnorm` (the reference's sqrt_simdot), optional person ids.

    dtype float32 -> exact-order fp32 scoring (bit-identical to the SSE code);
    dtype bfloat16 -> tcgen05 scoring (fp32 accumulate) for throughput."""

    def __init__(self, G: torch.Tensor, gnorm: torch.Tensor | None = None, person_ids: torch.Tensor | None = None,
                 idx_base: int = 0):
        L.require_cuda(G)
        if G.dim() != 2:
            raise ValueError("gallery must be (N, D)")
        self.G = G.contiguous()
        self.N, self.D = self.G.shape
        if gnorm is None:
            gnorm = row_norms(self.G) if G.dtype != torch.float32 else self._exact_norms()
        self.gnorm = gnorm.to(torch.float32).contiguous()
        self.person_ids = None if person_ids is None else person_ids.to(torch.int32).contiguous().to(G.device)
        self.idx_base = int(idx_base)

    def _exact_norms(self):
        # sqrt(simd_dot(fv, fv)) per row, the way registration computes sqrt_simdot (Feature.hpp:476)
        out = torch.empty(self.N, dtype=torch.float32, device=self.G.device)
        with torch.cuda.device(self.G.device):
            L.check(L.lib.efm_simd_norm_rows(L.ptr(self.G), self.N, self.D, L.ptr(out), L.stream()))
        return out

    # -- scores -------------------------------------------------------------------------------------------
    def scores(self, Q: torch.Tensor) -> torch.Tensor:
        """(nq, N) fp32 cosine scores, exact reference order (fp32 galleries only)."""
        if self.G.dtype != torch.float32:
            raise ValueError("scores(): exact scoring needs a float32 gallery")
        Q = Q.reshape(-1, self.D).to(torch.float32).contiguous()
        L.require_cuda(Q)
        out = torch.empty((Q.shape[0], self.N), dtype=torch.float32, device=self.G.device)
        with torch.cuda.device(self.G.device):
            L.check(L.lib.efm_gallery_scores_f32(L.ptr(self.G), L.ptr(self.gnorm), self.N, self.D, L.ptr(Q), Q.shape[0],
                                                 L.ptr(out), L.stream()))
        return out

    # -- top-K ----------------------------------------------------------------------------------------------
    def topk(self, Q: torch.Tensor, K: int = 5, sim_th: float = 0.0, positive_only: bool = False):
        """Per query the K best rows with score >= sim_th, ordered (score desc, index asc), NaN skipped.
        Returns (scores (nq,K) fp32, indices (nq,K) int32, -1 = unused slot)."""
        Q = Q.reshape(-1, self.D).to(self.G.dtype).contiguous()
        L.require_cuda(Q)
        nq = Q.shape[0]
        dev = self.G.device
        dt = L.dtype_code(self.G)
        out_s = torch.empty((nq, K), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, K), dtype=torch.int32, device=dev)
        ws = L.workspace(L.lib.efm_gallery_topk_workspace_bytes(self.N, self.D, nq, K, dt), dev)
        flags = L.GAL_POSITIVE_ONLY if positive_only else 0
        with torch.cuda.device(dev):
            L.check(L.lib.efm_gallery_topk(L.ptr(self.G), L.ptr(self.gnorm), self.N, self.D, L.ptr(Q), nq, K, float(sim_th),
                                           flags, dt, self.idx_base, L.ptr(out_s), L.ptr(out_i), L.ptr(ws), ws.numel(),
                                           L.stream()))
        return out_s, out_i

    def argmax(self, Q: torch.Tensor, sim_th: float):
        """Compare_Face_From_DB(FVV&, sim_th, fv) per query: index of the best row with sim >= sim_th and
        sim > 0 (maxsim starts at 0), lowest index on ties, -1 if none.  Returns (idx (nq,), sim (nq,))."""
        s, i = self.topk(Q, 1, sim_th, positive_only=True)
        return i.reshape(-1), s.reshape(-1)

    def identify(self, Q: torch.Tensor, sim_th: float):
        """Compare_Face_DB: persons are visited in ascending pid with a running strict maximum
        (Feature.hpp:799-801), so the winner is the person of the best row and, among rows with EXACTLY the best
        score, the lowest pid.  Returns (pid (nq,), best sim (nq,)); pid -1 if nobody passes.  Ties are resolved
        among the 16 best rows (more than 16 bit-identical best scores would need a wider list)."""
        if self.person_ids is None:
            raise ValueError("identify() needs person_ids")
        K = min(L.GAL_MAX_K, self.N)
        s, i = self.topk(Q, K, sim_th, positive_only=True)
        local = torch.where(i >= 0, i - self.idx_base, i).contiguous()
        pid = torch.empty_like(local)
        with torch.cuda.device(self.G.device):
            L.check(L.lib.efm_gather_pid(L.ptr(local), L.ptr(self.person_ids), local.numel(), L.ptr(pid), L.stream()))
        tied = (i >= 0) & (s == s[:, :1])
        big = torch.iinfo(torch.int32).max
        best = torch.where(tied, pid, torch.full_like(pid, big)).min(dim=1).values
        return torch.where(best == big, torch.full_like(best, -1), best), s[:, 0]


def merge_topk(part_scores: torch.Tensor, part_idx: torch.Tensor):
    """K-way merge of (parts, nq, K) partial lists (e.g. one per gallery shard) -> (nq, K)."""
    L.require_cuda(part_scores, part_idx)
    parts, nq, K = part_scores.shape
    ps = part_scores.to(torch.float32).contiguous(); pi = part_idx.to(torch.int32).contiguous()
    out_s = torch.empty((nq, K), dtype=torch.float32, device=ps.device)
    out_i = torch.empty((nq, K), dtype=torch.int32, device=ps.device)
    with torch.cuda.device(ps.device):
        L.check(L.lib.efm_topk_merge(L.ptr(ps), L.ptr(pi), parts, nq, K, L.ptr(out_s), L.ptr(out_i), L.stream()))
    return out_s, out_i


# ---- reference-named free functions -----------------------------------------------------------------------
def Compare_Face_From_DB(Mbs: Gallery, sim_th: float, Feature_Vector: torch.Tensor, Face_Maps_Size: int | None = None):
    """Both overloads: without Face_Maps_Size -> row index or -1 (Feature.hpp:345-392); with it -> the
    top-K list [(sim, row)] (Feature.hpp:295-343; the reference's names come from sqlite, rows here)."""
    if Face_Maps_Size is None:
        idx, _ = Mbs.argmax(Feature_Vector.reshape(1, -1), sim_th)
        return int(idx[0])
    s, i = Mbs.topk(Feature_Vector.reshape(1, -1), int(Face_Maps_Size), sim_th)
    return [(float(a), int(b)) for a, b in zip(s[0].tolist(), i[0].tolist()) if b >= 0]


def Compare_Face_Person(Mbs: Gallery, sim_th: float, Feature_Vector: torch.Tensor, maxsim: float):
    """Feature.hpp:763-790 -> (is_this_person, new maxsim)."""
    s, i = Mbs.topk(Feature_Vector.reshape(1, -1), 1, sim_th)
    if int(i[0, 0]) >= 0 and float(s[0, 0]) > maxsim:
        return True, float(s[0, 0])
    return False, float(maxsim)


def Compare_Face_DB(Mbs: Gallery, sim_th: float, Feature_Vector: torch.Tensor) -> int:
    """Feature.hpp:792-804 -> person id of the most similar registered person, -1 if none."""
    pid, _ = Mbs.identify(Feature_Vector.reshape(1, -1), sim_th)
    return int(pid[0])
