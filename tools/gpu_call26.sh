mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shards or c1_single or pipelined or ragged" > gpurun_out/r2/t30_shards.log 2>&1; echo shards rc=$?; tail -5 gpurun_out/r2/t30_shards.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/r2/multi_gpu_check_n2_v43.log 2>&1; echo check rc=$?; grep -E "rank|Error|error" gpurun_out/r2/multi_gpu_check_n2_v43.log | head -8
for sh in interleaved slabs; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 8 --warmup 4 --shards $sh > gpurun_out/r2/bench_c5_v43_n2_$sh.json 2> gpurun_out/r2/bench_c5_v43_n2_$sh.err; echo bench $sh rc=$?
tail -1 gpurun_out/r2/bench_c5_v43_n2_$sh.err
python - $sh <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r2/bench_c5_v43_n2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_update"],2), "e2e", round(d["e2e"]["ms_per_update"],2), d["multi_gpu_parity"], d["per_rank_trace_blend_ms"], d["parallelism"]["layout"])
PY
done
