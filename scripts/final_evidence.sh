#!/bin/bash
# Round-end evidence on one B200 box (run through gpurun from the repo root); everything lands in gpurun_out/r02f_*.
set -u
O=gpurun_out
mkdir -p $O
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02f_gpu_tests.log 2>&1; tail -3 $O/r02f_gpu_tests.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02f_smoke.log 2>&1; tail -2 $O/r02f_smoke.log
echo "== bench N=1"; timeout 1500 python bench.py --steps 3 --warmup 3 > $O/r02f_bench.log 2>&1; tail -1 $O/r02f_bench.log | cut -c1-600
echo "== bench decode"; timeout 900 python bench.py --config decode --steps 5 --warmup 3 --no-library-baseline > $O/r02f_bench_decode.log 2>&1; tail -1 $O/r02f_bench_decode.log | cut -c1-400
echo "== bench pair10"; timeout 900 python bench.py --config pair10 --steps 3 --warmup 3 --no-library-baseline > $O/r02f_bench_pair10.log 2>&1; tail -1 $O/r02f_bench_pair10.log | cut -c1-400
echo "== events"; timeout 600 python scripts/profile_unet_events.py --out $O/r02f_unet_events.txt > /dev/null 2>&1; head -3 $O/r02f_unet_events.txt
echo "== launch list (UNet)"; timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:"^(void )?<unnamed>" -c 1000 --csv --log-file $O/r02f_unet_launches.csv python scripts/bench_unet.py --no-graph --iters 1 --warmup 0 > $O/r02f_unet_launches.log 2>&1; wc -l $O/r02f_unet_launches.csv
echo "== launch list (VAE decode)"; timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:"^(void )?<unnamed>" -c 600 --csv --log-file $O/r02f_vae_launches.csv python scripts/bench_vae.py --no-graph --iters 0 --warmup 0 > $O/r02f_vae_launches.log 2>&1; wc -l $O/r02f_vae_launches.csv
for c in attn2560:tc_attn3 attnfusion:tc_attn3 attnwide:tc_attn_wide gn1280:gn_fused; do
  name=${c%%:*}; k=${c##*:}
  echo "== ncu $name"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $O/r02f_ncu_$name python scripts/prof_one.py $name > $O/r02f_ncu_$name.log 2>&1; tail -1 $O/r02f_ncu_$name.log | cut -c1-200
done
