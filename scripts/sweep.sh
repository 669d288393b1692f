b() { echo "== $*"; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-micro 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
b MVAE_HIPRI_LEVELS=0
b MVAE_HIPRI_LEVELS=1
b MVAE_HIPRI_LEVELS=2
b MVAE_HIPRI_LEVELS=2 MVAE_WGRAD_SMS=32
b MVAE_HIPRI_LEVELS=1 MVAE_WGRAD_SMS=148
b MVAE_HIPRI_LEVELS=0 MVAE_WGRAD_SMS=148
