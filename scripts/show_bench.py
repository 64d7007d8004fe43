import json,sys
for fn in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(fn) if l.startswith("{")][-1])
    except Exception as e:
        print(fn, "ERR", e); continue
    e=d["e2e"]
    print(fn.split("/")[-1], "value %.4g  ms %.3f  e2e_ms %.3f  list_ms %s  frac %.3f" % (d["value"], d["ms_per_step"], e.get("ms_per_step",0), e.get("ms_per_step_from_a_list_of_frame_arrays"), d["roofline"]["frac"]))
