cd $GRAFT_REPO_ROOT
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_backward.py tests/test_gpu_safe.py -x -q -m gpu 2>&1 | tail -3
echo "== bench"; python bench.py --no-cpu-baseline --no-api-line --steps 20 --warmup 3
