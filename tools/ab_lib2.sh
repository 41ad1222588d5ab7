# same-box A/B of library builds at a chosen step count:  bash tools/ab_lib2.sh "|--logits bf16" 100 libovdet_base.so libovdet.so
MODES=$1; STEPS=$2; shift; shift
IFS='|' read -ra MODE_LIST <<< "$MODES"
P=$PWD/$(ls -d real-time-*_b200)
for rep in 1 2; do
for mode in "${MODE_LIST[@]}"; do
for lib in "$@"; do
OVDET_LIB_PATH=$P/$lib timeout 180 python bench.py $mode --steps $STEPS --warmup 3 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '[$mode]', round(d['value']), round(d['stages_ms']['similarity'],4))"
done; done; done
