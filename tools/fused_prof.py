"""Per-phase cycle counters of the fused forward (k_gat_fused keeps them in its workspace): where a CTA's time goes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from magat_pathplanning_b200 import graphML

NAMES = ["scan+ximg", "wait(scan)", "lists", "score", "wait(score)", "attention", "wait(att)", "gather1", "wait(g1)",
         "gather2", "wait(g2)", "project", "mma:wait-operands", "mma:wait-acc", "epi:wait-acc", "tma:wait-stage"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4_n1000"
    w = dict(bench.WORKLOADS[name])
    if len(sys.argv) > 2:
        w["B"] = int(sys.argv[2])
    train = len(sys.argv) > 3 and sys.argv[3] == "train"
    dev = torch.device("cuda:0")
    layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
    layer.path = "fused"
    layer.max_degree = 16
    layer.fused_team = int(os.environ.get("MAGAT_TEAM", "8"))
    x = x_mem.permute(0, 2, 1)
    for _ in range(3):
        if train:
            layer.addGSO(S); y = layer(x.detach().requires_grad_(True))
        else:
            with torch.no_grad():
                layer.addGSO(S); y = layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        layer.addGSO(S); y = layer(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"event time of one forward: {e0.elapsed_time(e1):.3f} ms")
    ws = list(graphML._ws_cache.values())[0]
    off = (-ws.data_ptr()) % 1024
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    TEAM = int(os.environ.get("MAGAT_TEAM", "8"))
    T = min(sms // TEAM, w["B"])
    off_prof = (64 + T * 128 + 1023) // 1024 * 1024
    prof = ws[off + off_prof: off + off_prof + T * TEAM * 16 * 8].view(torch.int64).view(T * TEAM, 16).cpu().double()
    mhz = 1965.0
    per_inst = prof / (w["B"] / T) / mhz          # us per instance per CTA
    print(f"{name} B={w['B']} teams={T} {'train' if train else 'infer'}: per-instance us (mean over CTAs / max)")
    tot = 0
    for i, n in enumerate(NAMES):
        print(f"  {n:12s} {per_inst[:, i].mean():8.2f} {per_inst[:, i].max():8.2f}")
        tot += per_inst[:, i].mean() if i < 12 else 0
    print(f"  total        {tot:8.2f}   -> {tot * w['B'] / T / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
