# same-box A/B of library builds: bash tools/ab_epilogue.sh "v5 old" ; libovdet_<name>.so beside libovdet.so
P=$(ls -d real-time-*_b200)
for rep in 1 2; do
for mode in "" "--projected" "--logits bf16"; do
for lib in ${1:-old new}; do
OVDET_LIB_PATH=$PWD/$P/libovdet_$lib.so python bench.py $mode --steps 30 --warmup 3 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '$mode', round(d['value']), round(d['stages_ms']['similarity'],4))"
done; done; done
