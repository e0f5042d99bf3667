"""Print the bench JSON lines of one gpurun session directory in a compact form."""
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if d.get("impl") == "reference":
        print(f.split("/")[-1], "REFERENCE", round(d["value"], 1), d["unit"], d["cpu_baseline"]["cores"], "cores"); continue
    r = d["roofline"]
    print("%s: %.0f img/s  %.3f ms/step  launches/step %.0f  e2e %.0f img/s" % (
        d["config"]["workload"], d["value"], d["ms_per_step"], d["gpu_launches"] / d["steps"], d["e2e"]["value"]))
    print("   roofline: %s %s %s %.1f %s frac %.3f  %.1f us" % (r["kernel"], r["layer"], r["bound"], r["achieved"], r["unit"], r["frac"], r["us_per_launch"]))
    for k in d["kernels"][: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
        print("     %-16s %-45s %9.1f us  %.3f" % (k["call"], k["layer"], k["us"], k["share"]))
