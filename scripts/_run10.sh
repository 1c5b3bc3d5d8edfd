cd $GRAFT_REPO_ROOT
run() { echo "== $*"; python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
fb=d.get('fwd_bwd') or {'value':0,'e2e':{'value':0}}
print('fwd %.0f e2e %.0f fwd_bwd %.0f e2e %.0f api %s' % (d['value'], d['e2e']['value'], fb['value'], fb['e2e']['value'], (d.get('api_predict') or {}).get('us_per_rollout_step')))"; }
run --config inv_double_pendulum
PILCO_SETUP_FUSED=0 run --config inv_double_pendulum
run --config smgpr
