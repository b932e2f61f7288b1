set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_path.py > gpurun_out/sanitize_path_${tool}_r2x.log 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_path_${tool}_r2x.log
done
