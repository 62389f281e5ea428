import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],"value %.0f e2e %.0f ms/step %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]))
for k,v in d["kernels"].items(): print("   ",k,v["avg_us"],v["share"],v["frac"])
