"""Print the headline fields of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "spectra_per_s", "kernel_ms_per_step", "gpu_launches", "clocks", "n_gpus"):
    print(k, d.get(k))
print("roofline", d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"]["frac"])
print("e2e", d["e2e"]); print("cpu", d.get("cpu_baseline", {}).get("value"))
if "gradients" in d: print("grad", d["gradients"]["ms_per_step"], d["gradients"]["kernel_ms"])
if "plin" in d: print("plin", d["plin"]["ms"], d["plin"]["kmode_solves_per_s"])
