import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_face_recognition_performance_using_triplet_loss_b200 as P
N, D, nq = 125000, 512, int(os.environ.get("NQ", 8192))
G = torch.randn(N, D, device="cuda"); G = G / G.norm(dim=1, keepdim=True)
gal = P.Gallery(G.to(torch.bfloat16))
Q = (G[torch.randint(0, N, (nq,), device="cuda")] + 0.1 * torch.randn(nq, D, device="cuda")).to(torch.bfloat16)
for _ in range(3):
    s, i = gal.topk(Q, 5, 0.0)
torch.cuda.synchronize()
print("ok", s[0].tolist())
