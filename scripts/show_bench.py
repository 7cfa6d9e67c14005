"""Print the headline fields of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "spectra_per_s", "kernel_ms_per_step", "gpu_launches", "clocks", "n_gpus"):
    print(k, d.get(k))
print("roofline", d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"]["frac"])
print("e2e", d["e2e"]); print("cpu", d.get("cpu_baseline", {}).get("value"))
if "gradients" in d: print("grad", d["gradients"].get("ms_per_step"), d["gradients"].get("kernel_ms"), d["gradients"].get("error"))
if "plin" in d: print("plin", d["plin"].get("ms"), d["plin"].get("kmode_solves_per_s"), d["plin"].get("error"))
if "batch" in d: print("batch", d["batch"])
