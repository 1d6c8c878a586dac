"""Count the kernels in an `ncu --csv --metrics gpu__time_duration.sum` launch list by name (evidence that only frcnn:: kernels run)."""
import csv
import sys

names = {}
for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')):
    if r[0] == "ID":
        continue
    n = r[4].split("(")[0]
    names[n] = names.get(n, 0) + 1
tot = sum(names.values())
# ncu prints the kernel either with or without its namespace depending on the demangler mode: a FOREIGN kernel is one that carries
# another library's namespace (at::, cub::, thrust::, cutlass::, ...)
foreign = sum(c for n, c in names.items() if "::" in n.replace("frcnn::", ""))
print("launches: %d, of which from other libraries (at:: / cub:: / ...): %d" % (tot, foreign))
for n, c in sorted(names.items(), key=lambda kv: -kv[1]):
    print("  %5d  %s" % (c, n[:120]))
