mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/run_one.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_b.csv python profiles/run_one.py > gpurun_out/ncu1.log 2>&1
echo "ncu rc=$?"
