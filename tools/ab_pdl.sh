run() { # label, env, args
  lbl=$1; shift; envs=$1; shift
  env $envs timeout 300 python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lbl', d['config']['workload'], round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
}
for rep in 1 2; do
run pdl X=1 --workload edsr256_x4_b32_lr32 --math bf16
run nopdl SRB_NO_PDL=1 --workload edsr256_x4_b32_lr32 --math bf16
done
run pdl X=1 --workload espcn_x4_b128_lr64
run nopdl SRB_NO_PDL=1 --workload espcn_x4_b128_lr64
run pdl X=1 --workload edsr64_x4_b32_lr32
run nopdl SRB_NO_PDL=1 --workload edsr64_x4_b32_lr32
run pdl X=1 --workload srgan_x4_b16
run nopdl SRB_NO_PDL=1 --workload srgan_x4_b16
run pdl X=1 --workload vdsr_b64_128
