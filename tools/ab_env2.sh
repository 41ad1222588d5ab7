# same-box A/B over several environment settings:  bash tools/ab_env2.sh "A=1 B=2|A=0" "|--logits bf16"
# each setting is a space-separated list of VAR=value pairs; prints images/s and the similarity stage (ms)
SETS=$1; MODES=${2:-"|--logits bf16"}; STEPS=${3:-30}
IFS='|' read -ra SET_LIST <<< "$SETS"
IFS='|' read -ra MODE_LIST <<< "$MODES"
for rep in 1 2; do
for mode in "${MODE_LIST[@]}"; do
for st in "${SET_LIST[@]}"; do
env $st timeout 180 python bench.py $mode --steps $STEPS --warmup 3 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$st]', '[$mode]', round(d['value']), round(d['stages_ms']['similarity'],4))"
done; done; done
