import cProfile, pstats, sys, io
sys.argv = ["bench.py", "--config", "C4", "--steps", "30", "--warmup", "3", "--no-cpu-baseline"]
sys.path.insert(0, ".")
import bench
pr = cProfile.Profile()
pr.enable()
bench.main()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000], file=sys.stderr)
