cd /root/repo
bash tools/gpu_tests.sh -x -s > gpurun_out/gpu_tests_tail.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests_tail.log
python tools/gpu_wgrad_timing.py tc fp16 | tail -16
python tools/gpu_train_time.py > gpurun_out/train_time3.json 2> gpurun_out/train_time3.err; head -8 gpurun_out/train_time3.json
