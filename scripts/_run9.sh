cd $GRAFT_REPO_ROOT
echo "== tests"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
run() { echo "== $*"; python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
fb=d.get('fwd_bwd') or {'value':0,'e2e':{'value':0}}
print('fwd %.0f e2e %.0f fwd_bwd %.0f e2e %.0f api %s' % (d['value'], d['e2e']['value'], fb['value'], fb['e2e']['value'], (d.get('api_predict') or {}).get('us_per_rollout_step')))"; }
run
run --restarts 1 --no-backward
run --config inverted_pendulum --no-backward
run --config test_cascade --no-backward
