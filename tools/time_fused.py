"""Forward timing of the fused launch against the multi-launch path (CUDA events, device-resident inputs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4_n1000"
    w = dict(bench.WORKLOADS[name])
    if len(sys.argv) > 2:
        w["B"] = int(sys.argv[2])
    dev = torch.device("cuda:0")
    layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
    x = x_mem.permute(0, 2, 1)
    dy = dy_mem.permute(0, 2, 1)
    res = {}
    for path in ("tcgen05", "fused"):
        layer.path = path
        def fwd():
            with torch.no_grad():
                layer.addGSO(S)
                return layer(x)
        def train():
            for p in layer.parameters():
                p.grad = None
            xg = x.detach().requires_grad_(True)
            layer.addGSO(S)
            y = layer(xg)
            y.backward(dy)
        res[path] = {"fwd_ms": bench.timed(fwd, 10, 3, False), "train_ms": bench.timed(train, 10, 3, False)}
        if path == "fused":
            layer.max_degree = 16
            res["fused_trusted"] = {"fwd_ms": bench.timed(fwd, 10, 3, False), "train_ms": bench.timed(train, 10, 3, False)}
            layer.max_degree = None
    print(json.dumps({"workload": name, "B": w["B"], **res}))

if __name__ == "__main__":
    main()
