"""Developer timing of osc_batched_settle alone (graph prebuilt), CUDA events."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from oscillink_b200 import BatchedLattices

B = int(os.environ.get("B", "1440"))
mi_s = int(os.environ.get("MIS", "12")); mi_u = int(os.environ.get("MIU", "64"))
g = torch.Generator(device="cuda"); g.manual_seed(1)
Y = torch.randn((B, 1200, 384), generator=g, device="cuda")
psi = Y[:, :32].mean(1); psi = psi / psi.norm(dim=1, keepdim=True)
bl = BatchedLattices(Y, kneighbors=8); bl.set_query(psi)
for rep in range(3):
    bl.U = None
    out = bl.settle(max_iters=mi_s, tol=1e-3, receipt=True, ustar_max_iters=mi_u)
torch.cuda.synchronize()
print("debug", os.environ.get("OSC_BATCHED_DEBUG"), "B", B, "ms", bl.phase_ms()["batched_settle"],
      "iters", out["iters"].mean().item(), out["ustar_iters"].mean().item())
