mkdir -p gpurun_out
(timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_all.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|Error|^[a-z]" | cut -c1-200 | head -40) > gpurun_out/r2q_memcheck.txt; tail -2 gpurun_out/r2q_memcheck.txt
(timeout 250 compute-sanitizer --tool racecheck python tools/sanitize_all.py 2>&1 | grep -E "Error|RACECHECK|hazard|^[a-z]" | cut -c1-260 | sort | uniq -c | sort -rn | head -40) > gpurun_out/r2q_racecheck.txt; grep RACECHECK gpurun_out/r2q_racecheck.txt
