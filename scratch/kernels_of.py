import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fps %.1f ms/step %.4f F=%d e2e=%s e2e_wire=%s" % (d["value"], d["ms_per_step"], d["config"]["frames_per_step"], d.get("e2e") and round(d["e2e"]["value"], 1), d.get("e2e_wire") and round(d["e2e_wire"]["value"], 1)))
for k, v in d["kernels"].items():
    print("  %-14s %8.2f us  %8.1f GB/s" % (k, v["ms"] * 1000, v["GBps"]))
print("  path frac %.4f  top %s frac %.4f" % (d["roofline_path"]["frac"], d["roofline"]["kernel"], d["roofline"]["frac"]))
