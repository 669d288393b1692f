b() { echo "== $*"; env "$@" python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-micro --config $CFG 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
for CFG in cfg2 cfg3; do
echo "#### $CFG"
b MVAE_HIPRI_LEVELS=0
b MVAE_HIPRI_LEVELS=1
b MVAE_HIPRI_LEVELS=2
b MVAE_HIPRI_LEVELS=3
done
