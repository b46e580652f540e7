set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_final_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'lift32|tensor01|tensor32|relin12' --launch-skip 4 -c 4 -o gpurun_out/r02_fused_final -f python tools/prof.py mul 148 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/r02_final_gputests.txt; head -c 600 gpurun_out/r02_final_n1.json
