#!/bin/bash
# round-2 GPU call 9: programmatic dependent launch A/B (plain launches / PDL attribute only / PDL + early trigger) at
# batch 32 and 256, the full GPU test-suite under PDL, and the ncu DRAM-traffic capture of the two roofline kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c9_$name.log 2> gpurun_out/c9_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c9_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c9_$name.err
}
BARGS="--global-batch 32"
run_bench b32_pdl VG_PDL=1
run_bench b32_plain VG_PDL=0
run_bench b32_notrig VG_PDL=1 VG_LIB=$PWD/vae_gan_b200/lib/libvaegan_sm100_notrig.so
run_bench b32_pdl_2 VG_PDL=1
run_bench b32_plain_2 VG_PDL=0
BARGS="--global-batch 256"
run_bench b256_pdl VG_PDL=1
run_bench b256_plain VG_PDL=0
run_bench b256_notrig VG_PDL=1 VG_LIB=$PWD/vae_gan_b200/lib/libvaegan_sm100_notrig.so
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c9_tests.log
timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:"tc_conv_pair_kernel|bn_stream_kernel" -c 12 -o /tmp/c9_ncu_roofline python scripts/roofline_kernels.py > gpurun_out/c9_ncu_roofline.log 2>&1
ls -la /tmp/c9_ncu_roofline.ncu-rep
python scripts/ncu_traffic.py conv_128x128_fwd=/tmp/c9_ncu_roofline.ncu-rep:tc_conv_pair_kernel:64 bn_act_bwd_apply=/tmp/c9_ncu_roofline.ncu-rep:bn_stream_kernel:64
cp profiles/r2_ncu_traffic.json gpurun_out/c9_ncu_traffic.json
ncu -i /tmp/c9_ncu_roofline.ncu-rep --page raw --csv > gpurun_out/c9_ncu_roofline_raw.csv 2>/dev/null; ls -la gpurun_out/c9_ncu_roofline_raw.csv
