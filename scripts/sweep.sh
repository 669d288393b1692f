b() { echo "== $*"; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-micro --config $CFG 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
for CFG in cfg2 cfg3; do
echo "#### $CFG"
b MVAE_WGRAD_FLUSH_N=2
b MVAE_WGRAD_FLUSH_N=4
b MVAE_WGRAD_FLUSH_N=6
b MVAE_WGRAD_FLUSH_N=4 MVAE_WGRAD_SMS=32
done
