import sys, json; sys.path.insert(0,'/root/repo')
import bench
print(json.dumps(bench.bench_jpeg(), indent=1))
