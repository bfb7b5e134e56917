# Same-box A/B of the tensor-core weight-gradient kernel (chunk rows R, TMEM columns per CTA) on the training step:
#   tools/wg_ab.sh   (on the GPU box; prints step time and the wgrad_tc totals per configuration, sequential and concurrent)
for cfg in "256 512" "128 512" "128 256" "256 256"; do
  set -- $cfg
  echo "== R=$1 TMEM=$2"
  NVSE_CONCURRENT=0 NVSE_WG_R=$1 NVSE_WG_TMEM=$2 timeout 200 python tests/train_step_bench.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('step', round(d['ours_bf16_ms'],3), {k['kernel']: round(k['ms'],3) for k in d['profile'] if k['kernel'].startswith('wgrad_tc')})"
done
echo "== concurrent, R=128 TMEM=256"; NVSE_WG_R=128 NVSE_WG_TMEM=256 timeout 200 python tests/train_step_bench.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('step', round(d['ours_bf16_ms'],3))"
echo "== concurrent, default"; timeout 200 python tests/train_step_bench.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('step', round(d['ours_bf16_ms'],3))"
