import sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_umma import umma
for K in (32, 64, 128, 256, 512, 1024, 2048):
    g = torch.Generator().manual_seed(K)
    A = torch.randn(256, K, generator=g).cuda()
    B = (torch.randn(128, K, generator=g) / K ** 0.5).cuda()
    ref = A.double() @ B.double().t()
    out = umma(A, B)
    e = (out.double() - ref)
    f = ((A @ B.t()).double() - ref)
    # single-pass emulation: hi*hi only
    Ah = (A.view(torch.int32) & -8192).view(torch.float32); Bh = (B.view(torch.int32) & -8192).view(torch.float32)
    s = (Ah.double() @ Bh.double().t() - ref)
    print(f'K={K:5d} umma3 max {float(e.abs().max()/ref.abs().max()):.2e} mean-bias {float(e.mean()/ref.abs().mean()):+.2e} | torch fp32 {float(f.abs().max()/ref.abs().max()):.2e} | hi*hi exact {float(s.abs().max()/ref.abs().max()):.2e}')
# positive operands expose truncation bias
A = torch.rand(256, 1024).cuda(); B = torch.rand(128, 1024).cuda() / 32
ref = A.double() @ B.double().t(); out = umma(A, B)
print('positive operands K=1024: rel err max', float(((out.double()-ref)/ref).abs().max()), 'mean signed', float(((out.double()-ref)/ref).mean()))
