import sys, json
for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.0f MP/s  %.2f ms/step | e2e %.0f MP/s %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
        print({k: round(v["ms"], 3) for k, v in d["stages"].items()})
        print("pixels GB/s %.0f frac %.3f | dominant %s frac %.3f" % (d["roofline_pixels"]["achieved"], d["roofline_pixels"]["frac"], d["roofline"]["kernel"][:24], d["roofline"]["frac"]))
        print("clocks", d["clocks"], "cpu", d.get("cpu_baseline"))
    else:
        print(l.rstrip()[-400:])
