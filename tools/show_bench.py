"""Print the interesting parts of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1f %s  ms/step %.2f  e2e %s" % (d["value"], d["unit"], d["ms_per_step"], d.get("e2e")))
print("clocks", d.get("clocks"), "launches", d.get("gpu_launches"), "cpu", d.get("cpu_baseline"))
print("roofline", d.get("roofline"))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_launch"] * kv[1]["launches_per_step"]):
    print("  %-28s %8.4f ms x%-4g share %.4f  %s GB/s" % (k, v["ms_per_launch"], v["launches_per_step"], v["share_of_step"], v["algorithmic_gbs"]))
