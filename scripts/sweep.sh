b() { echo "== $*"; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-micro 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
b A=1
b MVAE_CONV_STAGES=3
b MVAE_CONV_STAGES=2
b MVAE_CONV_BUDGET=148
b MVAE_CONV_BUDGET=148 MVAE_CONV_STAGES=3
b MVAE_WG_STAGES=2
b MVAE_WG_STAGES=3
b MVAE_CONV_STAGES=3 MVAE_WG_STAGES=2
b MVAE_CONV_STAGES=3 MVAE_WG_STAGES=2 MVAE_WGRAD_SMS=148
b MVAE_WGRAD_SMS=148
