set -x
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/r5v_tests.txt
python bench.py > gpurun_out/r5v_bench.json 2> gpurun_out/r5v_bench.err
python bench.py --dtype bf16 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r5v_bench_bf16.json 2>> gpurun_out/r5v_bench.err
python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r5v_bench_dense.json 2>> gpurun_out/r5v_bench.err
python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/r5v_train.json 2>> gpurun_out/r5v_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5v_bench_reference.json 2>> gpurun_out/r5v_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r5v_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r5v_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"combine_runs|tc_layer_kernel|tc3_layer|grid_fill_planes|pack_maps|fcn_layer_kernel|vox_insert" -s 26 -c 16 -o gpurun_out/r5v_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --fusion-mode 2 > gpurun_out/r5v_ncu_top.log 2>&1
tail -3 gpurun_out/r5v_tests.txt
