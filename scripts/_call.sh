(timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2ad_pytest_gpu_n1_full.log 2>&1; echo pytest rc=$? >> gpurun_out/r2ad_pytest_gpu_n1_full.log); tail -14 gpurun_out/r2ad_pytest_gpu_n1_full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2ad_bench_cfg4_n1.json 2> gpurun_out/r2ad_bench_cfg4_n1.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/r2ad_bench_cfg4_n1.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['parity']['ok'], d.get('extra',{}).get('cfg2',{}).get('time_to_svd_s'))"
