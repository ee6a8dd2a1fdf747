#!/bin/bash
# round-2 GPU call 7: final single-GPU check - all tests, smoke, full bench line (with baselines), cfg2, ncu captures for the roofline traffic
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c7_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c7_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1; tail -1 gpurun_out/c7_smoke.log
timeout 900 python bench.py > gpurun_out/c7_bench.log 2> gpurun_out/c7_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads([l for l in open('gpurun_out/c7_bench.log') if l.startswith('{')][-1]);print('b256',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['vs_gpu_lib'], d['step_frac_of_ideal'])"
timeout 600 python bench.py --global-batch 64 --skip-cpu-baseline > gpurun_out/c7_bench_cfg2_b64.log 2> gpurun_out/c7_bench_cfg2.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c7_bench_cfg2_b64.log') if l.startswith('{')][-1]);print('b64',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['vs_gpu_lib'])"
timeout 600 python bench.py --global-batch 32 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c7_bench_b32.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/c7_bench_b32.log') if l.startswith('{')][-1]);print('b32',d['value'],d['ms_per_step'],d['launches_per_step'])"
timeout 900 python bench.py --workload cfg4 --skip-cpu-baseline --skip-lib-baseline --steps 10 > gpurun_out/c7_bench_cfg4_n1.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/c7_bench_cfg4_n1.log') if l.startswith('{')][-1]);print('cfg4 n1',d['value'],d['ms_per_step'],d['launches_per_step'], d['step_model_tflops_per_gpu'])"
timeout 600 python bench.py --loss-mode wgan_gp --optimizer rmsprop --skip-cpu-baseline --skip-lib-baseline --steps 10 > gpurun_out/c7_bench_wgan_gp.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/c7_bench_wgan_gp.log') if l.startswith('{')][-1]);print('wgan_gp',d['value'],d['ms_per_step'],d['launches_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_conv_pair_kernel|bn_stream_kernel" -c 40 -o gpurun_out/c7_ncu_roofline python scripts/roofline_kernels.py > gpurun_out/c7_ncu_roofline.log 2>&1
ls -la gpurun_out/c7_ncu_roofline.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c7_launches_b32.csv python scripts/profile_step.py 32 > gpurun_out/c7_ncu_b32.log 2>&1
python scripts/summarize_launches.py gpurun_out/c7_launches_b32.csv > gpurun_out/c7_launches_b32_summary.txt 2>&1; head -20 gpurun_out/c7_launches_b32_summary.txt
cuobjdump -sass vae_gan_b200/lib/libvaegan_sm100.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn > gpurun_out/c7_sass_histogram.txt; grep -E "UTCHMMA|UTMALDG|UBLKCP|LDTM|UTCBAR|SYNCS|RED|ATOM" gpurun_out/c7_sass_histogram.txt
