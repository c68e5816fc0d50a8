set -x
cd $GRAFT_REPO_ROOT
bash profiles/scripts/measure_round.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_grouped -c 1 -o gpurun_out/adam_full_final -f python profiles/adam_bench.py 1024 100 784 1024 1 > gpurun_out/ncu_adam_final.log 2>&1; tail -1 gpurun_out/ncu_adam_final.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_grouped -c 1 -o gpurun_out/fwd_full_final -f python profiles/linear_bench.py fwd 1024 100 784 1024 1 > gpurun_out/ncu_fwd_final.log 2>&1; tail -1 gpurun_out/ncu_fwd_final.log
for a in "fwd 1024 100 784" "fwd 512 100 1024" "fwd 784 200 512" "fwd 512 200 256" "fwd 784 100 512" "bwd 1024 100 784" "bwd 512 100 1024" "bwd 512 200 256" "bwd 784 100 512"; do python profiles/linear_bench.py $a 1024; done 2>&1 | grep -v "^+" > gpurun_out/linear_bench_final.log
python profiles/adam_bench.py 1024 100 784 1024 > gpurun_out/adam_bench_final.log; python profiles/adam_bench.py 784 200 512 1024 >> gpurun_out/adam_bench_final.log; python profiles/adam_bench.py 512 100 1024 1024 >> gpurun_out/adam_bench_final.log
cat gpurun_out/linear_bench_final.log gpurun_out/adam_bench_final.log
