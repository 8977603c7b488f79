import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f tok/s  %.3f ms/step | e2e %.0f tok/s | clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "frac", "share_of_step")} if d.get("roofline") else None)
ks = d["kernels"].get("classes", {})
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_per_step"])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("%-24s %7.3f ms %6.0f launches %8.2f us/launch  %s TF/s %s GB/s share %.3f" % (k, v["ms_per_step"], v["launches_per_step"], v["us_per_launch"], v["TFLOPps"] and round(v["TFLOPps"], 1), v["GBps"] and round(v["GBps"]), v["share"]))
print("sum kernels %.2f ms, launches/step %s, cost %s" % (d["kernels"].get("sum_kernel_ms_per_step", 0), d["kernels"].get("launches_per_step"), d["cost_first_last"]))
if "cpu_baseline" in d: print("cpu_baseline", d["cpu_baseline"])
