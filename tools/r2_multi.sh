# round-2 multi-GPU evidence run:  bash tools/r2_multi.sh N   (inside gpurun --gpus N)
N=$1; OUT=gpurun_out; P=29541; T=${2:-r2_final}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
run bench.py --gpus $N --steps 50 --warmup 5 > $OUT/${T}_bench_n${N}.json 2> $OUT/${T}_multi_n${N}.err
run bench.py --gpus $N --steps 30 --warmup 5 --classes 4800 --batch 16 --e2e-chunk 16 --no-cpu-baseline > $OUT/${T}_bench_n${N}_cfg4_c4800.json 2>> $OUT/${T}_multi_n${N}.err
run bench.py --gpus $N --steps 20 --warmup 5 --image-size 1280 --batch 64 --max-det 2048 --e2e-steps 5 --no-cpu-baseline > $OUT/${T}_bench_n${N}_cfg3_1280.json 2>> $OUT/${T}_multi_n${N}.err
run bench.py --gpus $N --steps 50 --warmup 5 --precision fp16 --classes 80 --batch 64 --no-cpu-baseline > $OUT/${T}_bench_n${N}_cfg1_fp16_c80.json 2>> $OUT/${T}_multi_n${N}.err
run bench.py --gpus $N --mode vocab-parallel --classes 76800 --batch 16 --steps 100 --warmup 10 > $OUT/${T}_vocab_parallel_n${N}_c76800_b16.json 2>> $OUT/${T}_multi_n${N}.err
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "across_gpus" 2>&1 | tail -3 > $OUT/${T}_test_vp_across_n${N}.log
tail -5 $OUT/${T}_multi_n${N}.err; cat $OUT/${T}_test_vp_across_n${N}.log
for f in $OUT/${T}_bench_n${N}.json $OUT/${T}_bench_n${N}_cfg4_c4800.json $OUT/${T}_bench_n${N}_cfg3_1280.json $OUT/${T}_bench_n${N}_cfg1_fp16_c80.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), d['e2e'] and round(d['e2e']['value']), d['clocks'] and d['clocks']['sm_mhz'])"; done
cut -c1-500 $OUT/${T}_vocab_parallel_n${N}_c76800_b16.json
