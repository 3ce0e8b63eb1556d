#!/bin/bash
# tools/gpu_session.sh — what one gpurun call runs on the B200 box; everything lands in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tests] [micro] [bench] [ncu] [sweep]'
set -u
mkdir -p gpurun_out
what="${*:-tests micro bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
for w in $what; do
case $w in
smoke)
  timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
tests)
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log ;;
micro)
  timeout 300 tiny_mc_b200/bin/tmc_microbench > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "micro rc=$?" ;;
bench)
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?" ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_walk -s 1 -c 1 -f -o gpurun_out/prof \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" ;;
ncucfg)
  for c in highalbedo finegrid; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_walk -s 1 -c 1 -f -o gpurun_out/prof_$c \
        python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$c.log 2>&1; echo "ncu $c rc=$?"
    timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"
  done ;;
report)
  timeout 900 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; echo "parity rc=$?"; python -c "
import json
d = json.load(open('gpurun_out/parity.json'))
for k, v in d.items(): print(k, {a: b for a, b in v.items() if a != 'z'})"
  timeout 900 python tools/soak.py > gpurun_out/soak.jsonl 2> gpurun_out/soak.err; echo "soak rc=$?"; cat gpurun_out/soak.jsonl ;;
sanitize)
  for t in memcheck racecheck initcheck; do timeout 900 compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitizer_$t.log 2>&1; echo "sanitizer $t rc=$?"; tail -2 gpurun_out/sanitizer_$t.log; done ;;
microncu)
  timeout 900 ncu --metrics sm__cycles_elapsed.avg,smsp__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
      --clock-control none --csv --log-file gpurun_out/micro_ncu.csv tiny_mc_b200/bin/tmc_microbench > gpurun_out/micro_under_ncu.jsonl 2>&1; echo "microncu rc=$?" ;;
r2a)
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "replay or other_optics or independent_of_split or single_photon or run_to_run" > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2a.log
  timeout 600 python tools/quick_bench.py > gpurun_out/quick.jsonl 2> gpurun_out/quick.err; echo "quick rc=$?"; cat gpurun_out/quick.jsonl
  for v in ppl1 oneatomic; do
    TMC_LIB=tiny_mc_b200/lib/exp/libtinymc_$v.so timeout 300 python tools/quick_bench.py default:0:0 default:768:1 default:1024:1 highalbedo:0:0 finegrid:0:0 >> gpurun_out/quick_variants.jsonl 2>> gpurun_out/quick.err; echo "variant $v rc=$?"
  done
  cat gpurun_out/quick_variants.jsonl ;;
scaleall)   # on an 8-GPU box: the driver's scaling run (N = 1, 2, 4, 8 back to back)
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "scale n=1 rc=$?"
  for n in 2 4 8; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "scale n=$n rc=$?"
  done
  python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/scale_n{n}.json"))
        print(n, f"value {d['value']:.4g} e2e {d['e2e']['value']:.4g} ms/step {d['ms_per_step']:.3f} hash {d['checks']['tally_hash']} lib {d['checks']['tally_hash_library_path']} launches {d['gpu_launches']}")
    except Exception as e:
        print(n, "failed", e)
PY
  ;;
scale)   # usage: gpurun --gpus N -- 'NGPU=N bash tools/gpu_session.sh scale'
  n=${NGPU:-2}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps ${STEPS:-5} --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "scale n=$n rc=$?"; cat gpurun_out/scale_n$n.json
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or batches" > gpurun_out/pytest_multigpu_n$n.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multigpu_n$n.log
  for prog in headless_config3; do
    TMC_TRACE=1 TMC_GPUS=$n timeout 300 tiny_mc_b200/bin/$prog > gpurun_out/${prog}_n$n.txt 2> gpurun_out/${prog}_n$n.err; echo "$prog rc=$?"; head -8 gpurun_out/${prog}_n$n.txt; tail -3 gpurun_out/${prog}_n$n.err
    TMC_TRACE=1 TMC_GPUS=$n TMC_JSON=gpurun_out/${prog}_n$n.json timeout 300 tiny_mc_b200/bin/$prog > gpurun_out/${prog}_json_n$n.txt 2> gpurun_out/${prog}_json_n$n.err; echo "$prog json rc=$?"; head -8 gpurun_out/${prog}_json_n$n.txt; tail -3 gpurun_out/${prog}_json_n$n.err
    TMC_TRACE=1 TMC_GPUS=$n TMC_NCCL=0 timeout 300 tiny_mc_b200/bin/$prog > gpurun_out/${prog}_hostsum_n$n.txt 2> gpurun_out/${prog}_hostsum_n$n.err; echo "$prog hostsum rc=$?"; head -8 gpurun_out/${prog}_hostsum_n$n.txt | tail -3; tail -2 gpurun_out/${prog}_hostsum_n$n.err
  done ;;
variants)   # quick_bench of the in-tree library and of every tiny_mc_b200/lib/exp/libtinymc_*.so
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "replay or other_optics or independent_of_split or single_photon" > gpurun_out/pytest_replay.log 2>&1; echo "pytest replay rc=$?"; tail -3 gpurun_out/pytest_replay.log
  rm -f gpurun_out/quick_variants.jsonl
  timeout 300 python tools/quick_bench.py ${PLANS:-default:0:0 default:768:1 default:1024:1 highalbedo:0:0:10:22 finegrid:0:0 finegrid:512:1} >> gpurun_out/quick_variants.jsonl 2>> gpurun_out/quick.err
  for f in tiny_mc_b200/lib/exp/libtinymc_*.so; do
    [ -e "$f" ] || continue
    TMC_LIB=$f timeout 300 python tools/quick_bench.py ${PLANS:-default:0:0 default:768:1 default:1024:1 highalbedo:0:0:10:22 finegrid:0:0 finegrid:512:1} >> gpurun_out/quick_variants.jsonl 2>> gpurun_out/quick.err; echo "variant $f rc=$?"
  done
  python - <<'PY'
import json
for l in open("gpurun_out/quick_variants.jsonl"):
    d = json.loads(l)
    if "error" in d: print(d); continue
    print(f"{d['lib'] or 'in-tree':28s} {d['config']:10s} block {d['block']:4d} x grid {d['grid']:3d} flush {d['flush']:4d}  {d['photons_per_s']:.4g} photons/s  {d['events_per_s']:.4g} events/s")
PY
  ;;
batches)   # the batched call (TMC_JSON) against the plain call, C host program, TMC_TRACE timings
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "batches or batch_means" > gpurun_out/pytest_batches.log 2>&1; echo "pytest batches rc=$?"; tail -3 gpurun_out/pytest_batches.log
  for prog in headless_config2 headless_config3; do
    TMC_TRACE=1 timeout 300 tiny_mc_b200/bin/$prog 2> gpurun_out/${prog}_plain.err | sed -n 8,9p; grep "tmc trace" gpurun_out/${prog}_plain.err | tail -1
    TMC_TRACE=1 TMC_JSON=gpurun_out/${prog}.json timeout 300 tiny_mc_b200/bin/$prog 2> gpurun_out/${prog}_json.err | sed -n 8,9p; grep "tmc trace" gpurun_out/${prog}_json.err | tail -1
    TMC_BATCH_STREAMS=1 TMC_TRACE=1 TMC_JSON=gpurun_out/${prog}.json timeout 300 tiny_mc_b200/bin/$prog 2> gpurun_out/${prog}_json1.err | sed -n 8,9p; grep "tmc trace" gpurun_out/${prog}_json1.err | tail -1
  done ;;
vtests)   # the exact-replay tests against variant libraries: VLIBS="name1 name2" (tiny_mc_b200/lib/exp/libtinymc_<name>.so)
  for v in ${VLIBS:-}; do
    TMC_LIB=tiny_mc_b200/lib/exp/libtinymc_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "replay or other_optics or independent_of_split or single_photon" > gpurun_out/pytest_replay_$v.log 2>&1; echo "pytest replay $v rc=$?"; tail -3 gpurun_out/pytest_replay_$v.log
  done ;;
vquick)   # quick_bench of every tiny_mc_b200/lib/exp/libtinymc_*.so (PLANS="config:block:per_sm ...")
  rm -f gpurun_out/quick_variants.jsonl
  for f in tiny_mc_b200/lib/exp/libtinymc_*.so; do
    TMC_LIB=$f timeout 300 python tools/quick_bench.py ${PLANS:-default:0:0 highalbedo:0:0:10:22 finegrid:0:0} >> gpurun_out/quick_variants.jsonl 2>> gpurun_out/quick.err; echo "variant $f rc=$?"
  done
  python - <<'PY'
import json
for l in open("gpurun_out/quick_variants.jsonl"):
    d = json.loads(l)
    if "error" in d: print(d); continue
    print(f"{d['lib'] or 'in-tree':28s} {d['config']:10s} block {d['block']:4d} x grid {d['grid']:3d} flush {d['flush']:4d}  {d['photons_per_s']:.4g} photons/s  {d['events_per_s']:.4g} events/s")
PY
  ;;
mixncu)   # the walk-mix ceiling micro-benchmark alone, with ncu's pipe counters
  timeout 300 ncu --metrics sm__cycles_elapsed.avg,smsp__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio \
      --clock-control none -k regex:k_walk_mix --csv --log-file gpurun_out/mix_ncu.csv tiny_mc_b200/bin/tmc_microbench > gpurun_out/mix_under_ncu.jsonl 2>&1; echo "mixncu rc=$?"; grep -E "k_walk_mix" gpurun_out/mix_ncu.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -14 ;;
quick)
  timeout 600 python tools/quick_bench.py > gpurun_out/quick.jsonl 2> gpurun_out/quick.err; echo "quick rc=$?"; cat gpurun_out/quick.jsonl ;;
sweep)
  timeout 900 python tools/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep rc=$?" ;;
esac
done
