import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, "resident %.3f ms  e2e %.3f ms  top-kernel launch %.3f ms"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["avg_launch_ms"]), d["roofline"]["kernel_time_share"])
    except Exception as e: print(f, "ERR", e)
