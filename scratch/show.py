import json,sys
for f in sys.argv[1:]:
    d=json.load(open(f))
    print(f, "value %.3g ms/step %.4f step_frac %.3f warm ms %.4f e2e %.3g launches %d" % (d["value"], d["ms_per_step"], d["step_roofline"]["frac"], d["l2_warm"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
    for k,v in d["kernels"].items(): print("   %-16s %7.1f us/launch x%.0f share %.2f" % (k, v["us_per_launch"], v["launches_per_step"], v["share"]))
