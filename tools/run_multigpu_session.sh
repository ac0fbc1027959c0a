#!/bin/bash
# One 8-GPU session: data-parallel bench lines (weak / strong scaling, peer vs NCCL SyncBN sums), the vertex-sharded full-sphere
# bench at Nside 512 / 1024, the multi-GPU parity tests.  Usage: gpurun --gpus 8 -- bash tools/run_multigpu_session.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
free -g | head -2
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
P=29600
run() { P=$((P+1)); timeout 600 $T --nproc-per-node $1 --master-port $P "${@:2}"; }
echo "== DP weak N=8 (peer all-reduce)"; run 8 bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/r02_bench_dp8.jsonl 2> $OUT/r02_bench_dp8.err; python tools/show_bench.py < $OUT/r02_bench_dp8.jsonl
echo "== DP weak N=8 (NCCL only)"; DSB200_PEER_ALLREDUCE=0 run 8 bench.py --gpus 8 --steps 20 --warmup 3 --no-parity --no-cpu-baseline 2>/dev/null > $OUT/r02_bench_dp8_nccl.jsonl; python tools/show_bench.py < $OUT/r02_bench_dp8_nccl.jsonl
echo "== DP strong N=8"; run 8 bench.py --gpus 8 --steps 20 --warmup 3 --scaling strong --no-parity --no-cpu-baseline 2>/dev/null > $OUT/r02_bench_dp8_strong.jsonl; python tools/show_bench.py < $OUT/r02_bench_dp8_strong.jsonl
echo "== DP weak N=4"; run 4 bench.py --gpus 4 --steps 20 --warmup 3 --no-parity --no-cpu-baseline 2>/dev/null > $OUT/r02_bench_dp4.jsonl; python tools/show_bench.py < $OUT/r02_bench_dp4.jsonl
echo "== DP weak N=2"; run 2 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --no-cpu-baseline 2>/dev/null > $OUT/r02_bench_dp2.jsonl; python tools/show_bench.py < $OUT/r02_bench_dp2.jsonl
echo "== DP strong N=4"; run 4 bench.py --gpus 4 --steps 20 --warmup 3 --scaling strong --no-parity --no-cpu-baseline 2>/dev/null > $OUT/r02_bench_dp4_strong.jsonl; python tools/show_bench.py < $OUT/r02_bench_dp4_strong.jsonl
: > $OUT/r02_sharded_bench.jsonl
echo "== sharded Nside 512: 1, 8 GPUs"
timeout 300 python tools/bench_sharded.py --nside 512 --graph --lmax 2>/dev/null | tail -1 | tee -a $OUT/r02_sharded_bench.jsonl | cut -c1-300
run 8 tools/bench_sharded.py --nside 512 --graph --lmax 2>/dev/null | tail -1 | tee -a $OUT/r02_sharded_bench.jsonl | cut -c1-300
DSB200_PEER_HALO=1 run 8 tools/bench_sharded.py --nside 512 --graph --lmax 2>/dev/null | tail -1 | cut -c1-300
MEM=$(free -g | awk '/Mem:/{print $2}')
if [ "$MEM" -ge 300 ]; then
  echo "== sharded Nside 1024: 8 GPUs, then 1 GPU"
  run 8 tools/bench_sharded.py --nside 1024 --graph --lmax 2>$OUT/r02_sharded_1024.err | tail -1 | tee -a $OUT/r02_sharded_bench.jsonl | cut -c1-300
  timeout 300 python tools/bench_sharded.py --nside 1024 --graph --lmax 2>/dev/null | tail -1 | tee -a $OUT/r02_sharded_bench.jsonl | cut -c1-300
else
  echo "host memory ${MEM} GB: skipping Nside 1024"
fi
echo "== multi-GPU parity tests"
timeout 900 python -m pytest tests/test_dp_parity.py tests/test_vertex_sharded.py -m gpu -q 2>&1 | tail -3
