import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.0f/s  %.4f ms/step  fwd+bwd %.0f/s  e2e %.0f/s  chain frac %.3f" % (d["value"], d["ms_per_step"], d["fwd_bwd"]["value"], d["e2e"]["value"], d["roofline"]["chain"]["frac"]))
    print("   " + "  ".join("%s %.1fus" % (s["kernel"], s["ms"]*1e3) for s in d["stages"]))
