(timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2af_pytest_multi_n2.log 2>&1; echo pytest rc=$? >> gpurun_out/r2af_pytest_multi_n2.log); tail -4 gpurun_out/r2af_pytest_multi_n2.log
SVDLAS2_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2af_bench_cfg4_n2.json 2> gpurun_out/r2af_bench_cfg4_n2.err; echo rc=$?
grep -E "^\[e2e\]|resident\]" gpurun_out/r2af_bench_cfg4_n2.err | tail -8 | cut -c1-330
grep -E "host pool" gpurun_out/r2af_bench_cfg4_n2.err | head -4 | cut -c1-260
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$(basename $d) numa_node=$(cat $d/numa_node)"; fi; done; ls /sys/devices/system/node/ | tr '\n' ' '; nproc
python -c "
import json; d=json.load(open('gpurun_out/r2af_bench_cfg4_n2.json')); print(d['ms_per_step'], d['value'], d['e2e']['time_to_svd_s'], d['roofline']['frac'], d['parity']['ok'])"
