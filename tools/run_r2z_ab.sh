# (the second build was copied to s2s-ismr-unet_b200/lib/ab/ for this run only; S2S_LIB makes _lib.py load it instead of the in-tree library)
# same-box A/B of two builds of the library (S2S_LIB): gconv look-ahead registers compact (ab/) vs one float4 per item (default)
set -x
C=s2s-ismr-unet_b200/lib/ab/libs2s_unet_compact.so
for rep in 1 2; do
  echo "compact $(S2S_LIB=$C python tools/profile_model.py --batch 16 --steps 400 | head -1 | cut -c100-180)"
  echo "default $(python tools/profile_model.py --batch 16 --steps 400 | head -1 | cut -c100-180)"
done
for v in compact default; do
  if [ $v = compact ]; then export S2S_LIB=$C; else unset S2S_LIB; fi
  python bench.py --steps 50 --warmup 5 --large-batch 0 --extras 0 --inference-c5 0 --concurrent-models 8 --no-cpu-baseline > gpurun_out/r2z_bench_$v.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench_$v.json').read().strip().splitlines()[-1]); print('$v bench', d['ms_per_step'], d['trial_batching']['value'])"
done
