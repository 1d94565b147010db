mkdir -p gpurun_out/san
for t in memcheck racecheck synccheck; do
  echo "== $t"; timeout 1200 compute-sanitizer --tool $t python tools/sanitize.py 2>&1 | grep -v "^CartPole\|^simple_spread\|^MountainCar\|^Acrobot\|^Pendulum" | tail -8
done > gpurun_out/san/sanitizer.txt 2>&1
cat gpurun_out/san/sanitizer.txt
