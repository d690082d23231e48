set -u
mkdir -p gpurun_out
bash tools/gpu/run.sh suite
bash tools/gpu/run.sh bench --steps 20 --warmup 3
for m in log_prob sample; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_chain_wino -s 3 -c 1 -o gpurun_out/prof_wino_$m -f python bench.py --mode $m --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/ncu_wino_$m.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/ncu_launches.log 2>&1
grep -E "nf_|td_" gpurun_out/launches.csv | tail -8 | cut -d, -f5,12- | cut -c1-160
ls -la gpurun_out/prof_wino_*.ncu-rep
