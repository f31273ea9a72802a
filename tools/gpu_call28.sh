mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/r2/multi_gpu_check_n8_v44.log 2>&1; echo check rc=$?; grep -E "rank|Error|error" gpurun_out/r2/multi_gpu_check_n8_v44.log | sort | head -18
for sh in interleaved slabs; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 4 --shards $sh > gpurun_out/r2/bench_c5_v44_n8_$sh.json 2> gpurun_out/r2/bench_c5_v44_n8_$sh.err; echo bench $sh rc=$?
tail -1 gpurun_out/r2/bench_c5_v44_n8_$sh.err
python - $sh <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r2/bench_c5_v44_n8_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_update"],2), "e2e", round(d["e2e"]["ms_per_update"],2), d["multi_gpu_parity"]["status"], d["per_rank_trace_blend_ms"], d["parallelism"]["layout"])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 8 --steps 20 --warmup 5 --workload c4 --shards slabs > gpurun_out/r2/bench_c4_v44_n8_slabs.json 2> gpurun_out/r2/bench_c4_v44_n8_slabs.err; echo c4 rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c4_v44_n8_slabs.json").read().strip().splitlines()[-1])
print("c4 slabs", round(d["ms_per_update"],3), "e2e", round(d["e2e"]["ms_per_update"],3), d["multi_gpu_parity"]["status"], d["per_rank_trace_blend_ms"])
PY
