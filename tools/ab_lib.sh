# same-box A/B of two builds of the library:  bash tools/ab_lib.sh "<modes separated by |>" libA.so libB.so
# (libovdet.so = the in-tree build; build the other one from a patched tree and copy it beside it)
MODES=${1:-"|--projected|--logits bf16"}; shift
IFS='|' read -ra MODE_LIST <<< "$MODES"
P=$PWD/$(ls -d real-time-*_b200)
for rep in 1 2; do
for mode in "${MODE_LIST[@]}"; do
for lib in "$@"; do
OVDET_LIB_PATH=$P/$lib timeout 120 python bench.py $mode --steps 30 --warmup 3 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '$mode', round(d['value']), round(d['stages_ms']['similarity'],4))"
done; done; done
