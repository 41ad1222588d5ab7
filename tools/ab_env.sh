# same-box A/B over an environment switch:  bash tools/ab_env.sh OVDET_EPI2 "0 1 3" "|--projected|--logits bf16"
VAR=$1; VALS=$2; MODES=${3:-"|--projected"}
IFS='|' read -ra MODE_LIST <<< "$MODES"
for rep in 1 2; do
for mode in "${MODE_LIST[@]}"; do
for v in $VALS; do
env $VAR=$v timeout 120 python bench.py $mode --steps 30 --warmup 3 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', '$mode', round(d['value']), round(d['stages_ms']['similarity'],4))"
done; done; done
