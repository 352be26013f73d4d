#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list + full capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag] [steps...]
# Every step runs under its own timeout and logs into gpurun_out/<tag>_*.log.
TAG=${1:-r1}
shift
STEPS=${@:-smoke tests bench launches ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
for s in $STEPS; do
  echo "=== $s ($(date +%T))"
  case $s in
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=8 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -22 $OUT/${TAG}_tests.log ;;
    testsall)
      timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -40 $OUT/${TAG}_tests.log ;;
    bench)
      timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err ;;
    benchref)
      timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>&1; echo "benchref rc=$?"; cat $OUT/${TAG}_bench_ref.json ;;
    tile)
      timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "tile or config3" > $OUT/${TAG}_tile.log 2>&1; echo "tile rc=$?"; tail -15 $OUT/${TAG}_tile.log
      timeout 600 python tools/bench_variants.py --only packets --quick > $OUT/${TAG}_packets.log 2>&1; echo "packets rc=$?"; cat $OUT/${TAG}_packets.log
      timeout 600 python tools/bench_variants.py --only perkey > $OUT/${TAG}_perkey.log 2>&1; echo "perkey rc=$?"; cat $OUT/${TAG}_perkey.log
      AGCM_PERKEY_TILE=0 timeout 600 python tools/bench_variants.py --only perkey > $OUT/${TAG}_perkey_notile.log 2>&1; echo "perkey (no tile) rc=$?"; cat $OUT/${TAG}_perkey_notile.log ;;
    perkey)
      timeout 600 python tools/bench_variants.py --only perkey > $OUT/${TAG}_perkey.log 2>&1; echo "perkey rc=$?"; cat $OUT/${TAG}_perkey.log
      AGCM_PERKEY_TILE=0 timeout 600 python tools/bench_variants.py --only perkey > $OUT/${TAG}_perkey_notile.log 2>&1; echo "perkey (no tile) rc=$?"; cat $OUT/${TAG}_perkey_notile.log ;;
    probe)
      for v in 256 512 1024 2048; do echo "warp min blocks $v"; for sz in "1024 16384" "16384 0" "8192 0" "4096 0" "16384 16384" "4096 65536"; do AGCM_WARP_MIN_BLOCKS=$v timeout 300 python tools/probe_aad_heavy.py 0 $sz 2>&1 | tail -1; done; done ;;
    hybrid)
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -o /tmp/hybrid_probe tools/hybrid_probe.cu -L aes-gcm-128-192-256-bits_b200 -laesgcm_b200 -Xlinker -rpath=$PWD/aes-gcm-128-192-256-bits_b200 > $OUT/${TAG}_hybrid.log 2>&1
      timeout 300 /tmp/hybrid_probe >> $OUT/${TAG}_hybrid.log 2>&1; echo "hybrid rc=$?"; cat $OUT/${TAG}_hybrid.log ;;
    ragged)
      timeout 600 python tools/bench_ragged.py > $OUT/${TAG}_ragged.jsonl 2>&1; echo "ragged rc=$?"; cat $OUT/${TAG}_ragged.jsonl
      AGCM_NO_LEN_SORT=1 timeout 600 python tools/bench_ragged.py > $OUT/${TAG}_ragged_nosort.jsonl 2>&1
      AGCM_NO_LEN_CLASSES=1 timeout 600 python tools/bench_ragged.py > $OUT/${TAG}_ragged_noclasses.jsonl 2>&1
      RAGGED_SLOTS_ONLY=1 AGCM_SLOTS_LANES=2048 timeout 600 python tools/bench_ragged.py > $OUT/${TAG}_ragged_gather.jsonl 2>&1 ;;
    e2esizes)
      timeout 600 python tools/bench_e2e.py 32:0 32:1024 32:2048 16:1024 64:1024 > $OUT/${TAG}_e2e_sizes.jsonl 2>&1; echo "e2esizes rc=$?"; cat $OUT/${TAG}_e2e_sizes.jsonl
      for k in none tiny d2d kstream indep; do timeout 120 python tools/pipeline_timeline.py 512 16 $k q; done > $OUT/${TAG}_timeline.txt 2>&1; cat $OUT/${TAG}_timeline.txt ;;
    sweep5)
      timeout 1500 python tools/sweep_config5.py --out $OUT/${TAG}_config5.json > $OUT/${TAG}_config5.log 2>&1; echo "sweep5 rc=$?"; tail -75 $OUT/${TAG}_config5.log ;;
    variants)
      timeout 900 python tools/bench_variants.py > $OUT/${TAG}_variants.log 2>&1; echo "variants rc=$?"; cat $OUT/${TAG}_variants.log ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_launches_run.log 2>&1; echo "launches rc=$?"; tail -3 $OUT/${TAG}_launches_run.log ;;
    ncu)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -f -o $OUT/${TAG}_prof \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu_run.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu_run.log ;;
    ncutile)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_batch_tile -s 2 -c 1 -f -o $OUT/${TAG}_prof_tile \
        python tools/bench_variants.py --only packets --quick > $OUT/${TAG}_ncutile_run.log 2>&1; echo "ncutile rc=$?"; tail -3 $OUT/${TAG}_ncutile_run.log ;;
    ncuwarp)
      timeout 900 ncu --set full --clock-control none -k regex:k_batch_warp -s 1 -c 1 -f -o $OUT/${TAG}_prof_warp \
        python tools/probe_aad_heavy.py > $OUT/${TAG}_ncuwarp_run.log 2>&1; echo "ncuwarp rc=$?"; tail -3 $OUT/${TAG}_ncuwarp_run.log
      timeout 900 ncu --set full --clock-control none -k regex:k_batch_warp -s 1 -c 1 -f -o $OUT/${TAG}_prof_warp2 \
        python tools/probe_aad_heavy.py 0 262144 4194304 > $OUT/${TAG}_ncuwarp2_run.log 2>&1; echo "ncuwarp2 rc=$?"; tail -3 $OUT/${TAG}_ncuwarp2_run.log ;;
    ncubatch)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_batch -s 2 -c 1 -f -o $OUT/${TAG}_prof_batch \
        python tools/bench_variants.py --only packets --quick > $OUT/${TAG}_ncubatch_run.log 2>&1; echo "ncubatch rc=$?"; tail -3 $OUT/${TAG}_ncubatch_run.log ;;
    ncuperkey)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_batch_perkey -s 1 -c 1 -f -o $OUT/${TAG}_prof_perkey \
        python tools/bench_variants.py --only perkey --quick > $OUT/${TAG}_ncuperkey_run.log 2>&1; echo "ncuperkey rc=$?"; tail -3 $OUT/${TAG}_ncuperkey_run.log ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitize.log 2>&1; echo "sanitize rc=$?"; tail -8 $OUT/${TAG}_sanitize.log ;;
    scale4|scale8)
      N=${s#scale}
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
        bench.py --gpus $N --steps 30 --warmup 5 > $OUT/${TAG}_scale$N.json 2> $OUT/${TAG}_scale$N.err; echo "scale$N rc=$?"; cat $OUT/${TAG}_scale$N.json; tail -3 $OUT/${TAG}_scale$N.err ;;
    scale8nccl)
      AGCM_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 \
        bench.py --gpus 8 --steps 30 --warmup 5 > $OUT/${TAG}_scale8nccl.json 2> $OUT/${TAG}_scale8nccl.err; echo "scale8nccl rc=$?"; cat $OUT/${TAG}_scale8nccl.json; tail -3 $OUT/${TAG}_scale8nccl.err ;;
    refN8)
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 \
        bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $OUT/${TAG}_ref8.json 2> $OUT/${TAG}_ref8.err; echo "ref8 rc=$?"; cat $OUT/${TAG}_ref8.json; tail -3 $OUT/${TAG}_ref8.err ;;
    sanitize2)
      timeout 1700 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x --timeout 1600 -k "tile or bulk_aad or peer_exchange_single or fails_closed or verified or long_iv_shard or few_long or split_over_ranks" > $OUT/${TAG}_sanitize2.log 2>&1; echo "sanitize2 rc=$?"; tail -12 $OUT/${TAG}_sanitize2.log ;;
    sanitizeall)
      timeout 2400 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q --timeout 2300 -k "not (8GiB or maximum_length or full_size or two_gpus or plain_c_client)" > $OUT/${TAG}_sanitizeall.log 2>&1; echo "sanitizeall rc=$?"; tail -12 $OUT/${TAG}_sanitizeall.log ;;
    racecheck2)
      timeout 1700 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x --timeout 1600 -k "tile or bulk_aad or few_long" > $OUT/${TAG}_racecheck2.log 2>&1; echo "racecheck2 rc=$?"; tail -12 $OUT/${TAG}_racecheck2.log ;;
    racecheck)
      timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 $OUT/${TAG}_racecheck.log ;;
    pcie8|pcie4|pcie2)
      N=${s#pcie}
      nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N \
        tools/bench_pcie_multi.py > $OUT/${TAG}_pcie$N.json 2> $OUT/${TAG}_pcie$N.err; echo "pcie$N rc=$?"; cat $OUT/${TAG}_pcie$N.json; tail -3 $OUT/${TAG}_pcie$N.err ;;
    scale2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus 2 --steps 30 --warmup 5 > $OUT/${TAG}_scale2.json 2> $OUT/${TAG}_scale2.err; echo "scale2 rc=$?"; cat $OUT/${TAG}_scale2.json; tail -5 $OUT/${TAG}_scale2.err ;;
  esac
done
# full ncu reports are 20-40 MB each and gpurun copies back at most 64 MiB: keep their raw pages and summaries
for r in $OUT/${TAG}_prof*.ncu-rep; do
  [ -f "$r" ] || continue
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  case $b in
    *_prof) k=k_stream ;; *_prof_tile) k=k_batch_tile ;; *_prof_warp|*_prof_warp2) k=k_batch_warp ;; *_prof_batch) k=k_batch ;; *_prof_perkey) k=k_batch_perkey ;; *) k=k_ ;;
  esac
  python tools/ncu_summary.py $r $k > $b.summary.md 2>/dev/null
  [ "$k" = k_stream ] && python tools/ncu_json.py $r 1073741824 "gpurun_out/$(basename $b).raw.csv (ncu --set full --clock-control none, round 2)" > $b.json 2>/dev/null
  [ -n "$KEEP_REP" ] || rm -f $r
done
echo "=== done ($(date +%T))"
