#!/bin/bash
# round-end verification on one B200 (run through gpurun): tests, smoke, bench (both arms), ncu launch list + captures,
# compute-sanitizer.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/final_pytest.log
python __graft_entry__.py smoke > $O/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/final_smoke.log
python bench.py --impl reference > $O/final_ref.json 2> $O/final_ref.err; echo "reference arm rc=$?"
python bench.py > $O/final_n1.json 2> $O/final_n1.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras --skip-dropin > $O/final_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
for k in score_gemm refine_cert rescan_kernel upsample_hblur; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/prof_$k \
      python scripts/profile_target.py score 16 > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > $O/final_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > $O/final_racecheck.log 2>&1; echo "racecheck rc=$?"
for f in $O/final_memcheck.log $O/final_racecheck.log; do tail -n 2 $f; done
