import json,sys
d=json.load(open(sys.argv[1]))
print("train ms", round(d["ms_per_step"],3), "fwd ms", round(d["fwd"]["ms_per_step"],3), "e2e ms", d["e2e"] and round(d["e2e"]["ms_per_step"],2), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
for k in d["roofline"]["kernels"]: print("  %-40s x%d %8.3f ms" % (k["kernel"], k["launches_per_step"], k["ms_per_step"]))
print(" train step kernels:")
for k in d.get("train_kernels", []): print("  %-40s x%d %8.3f ms" % (k["kernel"], k["launches_per_step"], k["ms_per_step"]))
print("frac", d["roofline"]["frac"], "fwd_kernel_ms", d["roofline"]["fwd_kernel_ms_per_step"], d["clocks"], d["gpu_launches_per_step"])

if d.get("positions_input"): print("positions input:", {k: (round(v["ms_per_step"], 3) if isinstance(v, dict) and "ms_per_step" in v else v) for k, v in d["positions_input"].items() if k != "note"})
if d.get("cuda_graph"): print("cuda graph:", {k: (round(v["ms_per_step"], 4) if isinstance(v, dict) and "ms_per_step" in v else v) for k, v in d["cuda_graph"].items() if k != "note"})
