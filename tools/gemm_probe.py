import os, sys, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import ops
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
out = {}
for m, n, k, tn in ((490000, 768, 192, 128), (490000, 192, 192, 64), (1400, 256, 2112, 32), (8192, 8192, 8192, 128)):
    x = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda'); y = torch.empty(m, n, device='cuda')
    out[f'{m}x{n}x{k}/bn{tn}'] = round(t(lambda: ops.linear(x, w, out=y, tile_n=tn)), 4)
print(os.environ.get('ABX_GEMM_DEBUG_SKIP', '0'), json.dumps(out))
