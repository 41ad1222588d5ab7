# round-2 single-GPU evidence run at the final build:  bash tools/r2_final.sh   (inside gpurun, one GPU)
OUT=gpurun_out; T=r2f
b() { name=$1; shift; timeout 600 python bench.py "$@" > $OUT/${T}_bench_n1_$name.json 2>> $OUT/${T}.err; tail -c 100 $OUT/${T}_bench_n1_$name.json | head -c 0; python - "$OUT/${T}_bench_n1_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[2], round(d["value"]), "img/s", round(d["ms_per_step"], 4), "ms/step", "sim", round(d["stages_ms"]["similarity"], 4) if d.get("stages_ms") else None,
          r.get("bound"), round(r.get("frac", 0), 3), "e2e", d["e2e"] and round(d["e2e"]["value"]), "clk", d["clocks"] and d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b default --steps 20 --warmup 3
b 100 --steps 100 --warmup 5
b logits_bf16 --steps 50 --warmup 3 --logits bf16
b fp16 --steps 50 --warmup 3 --precision fp16
b per_image_text --steps 20 --warmup 3 --per-image-text
b projected --steps 50 --warmup 3 --projected
b bf16_activations --steps 50 --warmup 3 --input-dtype bf16
b cfg1_fp16_c80 --steps 50 --warmup 3 --precision fp16 --classes 80 --batch 64
b cfg1_fp32_c80 --steps 50 --warmup 3 --precision fp32 --classes 80 --batch 64
b cfg3_1280 --steps 20 --warmup 3 --image-size 1280 --batch 64 --max-det 2048 --e2e-steps 5
b cfg4_c4800 --steps 30 --warmup 3 --classes 4800 --batch 16 --e2e-chunk 16
b fp32_c1203_two_kernel --steps 10 --warmup 3 --precision fp32 --e2e-steps 2 --no-cpu-baseline
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${T}_bench_reference_arm.json 2>> $OUT/${T}.err; cut -c1-200 $OUT/${T}_bench_reference_arm.json
timeout 300 python tools/latency_probe.py > $OUT/${T}_latency_probe.txt 2>> $OUT/${T}.err; tail -5 $OUT/${T}_latency_probe.txt
timeout 120 python tools/bench_attention.py > $OUT/${T}_attention.json 2>> $OUT/${T}.err; cat $OUT/${T}_attention.json
tail -3 $OUT/${T}.err
