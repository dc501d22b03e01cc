set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1 > gpurun_out/r2a_env.txt
nproc >> gpurun_out/r2a_env.txt; free -g | head -2 >> gpurun_out/r2a_env.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt
cat gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
python tools/decode_probe.py 29 > gpurun_out/r2a_probe.txt 2>&1
cat gpurun_out/r2a_probe.txt
