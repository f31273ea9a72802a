mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/r2/multi_gpu_check_n2.log 2>&1; echo check rc=$?; grep -E "rank|Error" gpurun_out/r2/multi_gpu_check_n2.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2/bench_c5_v27_n2.json 2> gpurun_out/r2/bench_c5_v27_n2.err; echo bench rc=$?
tail -3 gpurun_out/r2/bench_c5_v27_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v27_n2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],2), d["stage_ms"]["march"], d["e2e"]["ms_per_update"], d["multi_gpu_parity"], d["config"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2/bench_ref_n2.json 2> gpurun_out/r2/bench_ref_n2.err; echo ref rc=$?; cut -c1-400 gpurun_out/r2/bench_ref_n2.json
