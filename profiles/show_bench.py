"""Print the judged fields of a bench.py JSON line: python profiles/show_bench.py gpurun_out/x.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
j = lambda o, n=900: json.dumps(o)[:n]   # noqa: E731
print("n_gpus", d.get("n_gpus"), "value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 3), "dtype", d["dtype"],
      "launches", d.get("gpu_launches"), "fallbacks", d.get("status_fallback_halfspaces"))
print("roofline", j({k: v for k, v in d["roofline"].items() if k not in ("peak_source",)}))
print("clocks", j(d.get("clocks")))
print("e2e", j({k: v for k, v in d["e2e"].items() if k != "note"}))
print("cpu_baseline", j(d.get("cpu_baseline"), 200), "parity", j(d.get("parity_spot_check"), 300))
for k in ("generated", "config5", "f64_inputs", "strong_scaling", "small_n"):
    print(k, j(d.get(k), 1600))
