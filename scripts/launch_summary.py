"""Prints the last `per_step` launches of an ncu gpu__time_duration launch list (CSV) with their durations."""
import csv
import re
import sys

path, per_step = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
step = rows[-per_step:]
total = 0
for r in step:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("unnamed>::", "").replace("eml::", "")
    print(f"{r[0]:>4} {name[:60]:60} {r[8]:>16} {int(r[-1]) / 1e3:8.2f} us")
    total += int(r[-1])
print(f"launches {len(step)}  total {total / 1e3:.1f} us")
