"""Small driver for ncu: a few rollout forward/backward passes at the benchmark batch (B=256), short sequence."""
import sys, torch
sys.path.insert(0, '.')
from humor_b200 import synth
from humor_b200.humor_model import HumorModel
from tests.test_gpu_kernels import make_state
B, S = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 6
m = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
m.load_state_dict(synth.make_humor_state_dict()); m = m.cuda().eval()
x0 = torch.tensor(make_state(B, 1)).cuda().requires_grad_(True)
z = (torch.randn(B, S, 48) * 0.5).cuda().requires_grad_(True)
for _ in range(2):
    w, p = m.roll_out_raw(x0, z, True)
    (w.sum() + p.sum()).backward()
torch.cuda.synchronize()
