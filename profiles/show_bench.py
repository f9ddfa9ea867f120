import json,sys
d=json.loads(sys.stdin.read())
print("value %.1f ms %.4f eager %.4f core %.4f"%(d["value"], d["ms_per_step"], d["eager_launch"]["ms_per_step"], d["core"]["ms_per_step"]))
for k,v in d["kernels"].items(): print("  %-24s %.4f"%(k, v["ms"]))
