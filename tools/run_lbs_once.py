"""Driver for ncu: dense LBS forward at the benchmark size (N = 256 x 60 frames)."""
import sys, torch
sys.path.insert(0, '.')
from humor_b200 import synth
from humor_b200.body_model import BodyModel, lbs
B, T = 256, 60
bm = BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=B * T, use_vtx_selector=True).to('cuda')
g = torch.Generator().manual_seed(0)
N = B * T
ro = (torch.randn(N, 3, generator=g) * 0.5).cuda(); pb = (torch.randn(N, 63, generator=g) * 0.3).cuda()
be = (torch.randn(B, 16, generator=g) * 0.5).cuda(); tr = torch.randn(N, 3, generator=g).cuda()
with torch.no_grad():
    for _ in range(2):
        lbs(bm.lbs_model, ro, pb, be, tr, T, None, True, False, 73)
torch.cuda.synchronize()
