run() { env "$@" timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline > /tmp/b.log 2>/tmp/b.err; python - "$*" <<PY
import json,sys
d=json.loads([l for l in open("/tmp/b.log") if l.startswith("{")][-1])
print(sys.argv[1], d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"])
PY
}
run A=1
run SRVP_WGRAD_SMEM_KB=227
run SRVP_WGRAD_HOLD_RES=32
run SRVP_WGRAD_HELD_CTAS=148
run SRVP_WGRAD_HELD_CTAS=96
run A=2
