# usage: bash tools/run_final_n1.sh <tag>  -- the round's 1-GPU evidence: GPU suite, smoke, the three bench workloads, the reference arm, ncu lists
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/${tag}_n1_bench_final.json 2> gpurun_out/${tag}_n1_bench_final.err; python tools/show.py gpurun_out/${tag}_n1_bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_n1_ref_final.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_n1_ref_final.json
timeout 400 python bench.py --workload M > gpurun_out/${tag}_m_bench_n1.json 2> gpurun_out/${tag}_m.err; python tools/show.py gpurun_out/${tag}_m_bench_n1.json
timeout 400 python bench.py --workload nets > gpurun_out/${tag}_nets_bench.json 2> gpurun_out/${tag}_nets.err; python tools/show.py gpurun_out/${tag}_nets_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 80 -c 60 --csv --log-file gpurun_out/${tag}_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustained-steps 0 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sgm_v2_kernel -s 2 -c 2 -o gpurun_out/${tag}_prof_v python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sustained-steps 0 > /dev/null 2>&1
ls -la gpurun_out/${tag}_prof_v.ncu-rep gpurun_out/${tag}_step_metrics.csv
