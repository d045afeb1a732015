"""GPU check of the tcgen05 layer kernel against the exact-fp32 SIMT kernel and an fp64 torch evaluation."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import _lib
from mvxnet_makise_b200 import modules as M

torch.manual_seed(0)
dev = 'cuda'
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def ref64(x, w, b, eps=1e-6):
    y = torch.relu(x.double() @ w.double().t() + b.double())
    m = y.mean(0)
    v = y.var(0, unbiased=False)
    return (y - m) / torch.sqrt(v + eps)


def err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


for (R, cin, cout) in [(1000, 768, 768), (5003, 768, 128), (777, 128, 128), (256, 768, 768), (100000, 768, 768)]:
    x = torch.randn(1, R // 7 if False else R, 1, cin, device=dev)
    x[:, R // 2:] *= 0.01
    fcn = M.FCN(cin, cout).to(dev)
    with torch.no_grad():
        _lib.set_gemm_mode(0)
        y0 = fcn(x)
        torch.cuda.synchronize()
        _lib.set_gemm_mode(MODE)
        t0 = time.perf_counter()
        y1 = fcn(x)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        r = ref64(x.reshape(-1, cin), fcn.fc.weight, fcn.fc.bias)
    print(f'R={R} {cin}->{cout}: tc-vs-simt {err(y1, y0):.3e}  simt-vs-f64 {err(y0.reshape(-1, cout), r):.3e}  tc-vs-f64 {err(y1.reshape(-1, cout), r):.3e}  nan={bool(torch.isnan(y1).any())}  ({(t1 - t0) * 1e3:.2f} ms incl. host)', flush=True)
print('tc_check done')
