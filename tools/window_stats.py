import sys, torch, math
sys.path.insert(0, ".")
from dv_matcher_b200 import synthetic
d = synthetic.make_batch(1, 50000, 50000, regime="structured")
X = d["feat1"][0].cuda(); Y = d["feat2"][0].cuda()
rows = torch.randperm(50000)[:2048].cuda()
D = torch.cdist(X[rows], Y)
dmin = D.min(1).values
cut = min(32.0, math.log(50000) + 11.6) / 100.0
inwin = (D < (dmin[:, None] + cut)).sum(1).float()
k16 = D.kthvalue(16, dim=1).values
print("cut", cut, "window count mean/median/max", inwin.mean().item(), inwin.median().item(), inwin.max().item())
print("dmin mean", dmin.mean().item(), "d16 - dmin mean", (k16 - dmin).mean().item())
# chunks (16 columns) containing a window entry, per row
ch = (D < (dmin[:, None] + cut)).view(2048, -1, 16).any(-1).sum(1).float()
print("chunks with window entries per row mean", ch.mean().item(), "of", D.shape[1] // 16)
ch16 = (D < k16[:, None] * 1.0000001).view(2048, -1, 16).any(-1).sum(1).float()
print("chunks with top16 entries", ch16.mean().item())
