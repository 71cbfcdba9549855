import sys, os, torch, time
sys.path.insert(0, '/root/repo')
import bench
from magat_pathplanning_b200 import GraphFilterBatch, _cabi
dev = torch.device('cuda:0')
B, N, G, F, K = 128, 1000, 128, 128, 3
S = bench.synth_gso(B, N, 200, dev, torch.Generator(device=dev).manual_seed(1))
layer = GraphFilterBatch(G, F, K, 1, True).to(dev)
x = torch.relu(torch.randn(B, N, G, device=dev)).permute(0, 2, 1)
dy = torch.randn(B, N, F, device=dev).permute(0, 2, 1)
def step():
    for p in layer.parameters(): p.grad = None
    xg = x.detach().requires_grad_(True)
    layer.addGSO(S); layer(xg).backward(dy)
for _ in range(3): step()
torch.cuda.synchronize()
L = _cabi.lib(); L.magat_profile_enable(1); step(); torch.cuda.synchronize()
rec = _cabi.profile_collect(); L.magat_profile_enable(0)
print('GraphFilterBatch total %.3f ms ' % sum(r[2] for r in rec) + ' '.join('%s=%.3f' % (r[0], r[2]) for r in sorted(rec, key=lambda r: -r[2])[:6]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize(); print('wall ms/step', e0.elapsed_time(e1) / 10)
