b() { echo "== $*"; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-micro --config $CFG 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
for CFG in cfg2 cfg3 cfg4; do
echo "#### $CFG"
b A=1
b MVAE_NO_DEFER_WGRAD=1
done
