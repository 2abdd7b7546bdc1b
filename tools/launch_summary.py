"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last step only).

    python tools/launch_summary.py gpurun_out/launches_r1.csv [steps_in_capture]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    data = [row for row in r]
    # a step ends with the fused SGD kernel: use it as the delimiter
    ends = [i for i, row in enumerate(data) if 'sgd_kernel' in row[idx['Kernel Name']]]
    if len(ends) >= 2:
        last = data[ends[-2] + 1: ends[-1] + 1]
    else:
        per = len(data) // steps
        last = data[-per:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in last:
        name = row[idx['Kernel Name']]
        short = re.sub(r'^void ', '', name)
        short = re.sub(r'\(.*', '', short)[:72]
        v = float(row[idx['Metric Value']].replace(',', ''))
        agg[short][0] += 1
        agg[short][1] += v
        tot += v
    print(f"launches in step: {len(last)}   sum of kernel durations: {tot / 1e6:.3f} ms")
    print(f"{'share':>7} {'ms':>9} {'n':>5}  kernel")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / tot * 100:6.2f}% {v / 1e6:9.3f} {c:5d}  {k}")


if __name__ == '__main__':
    try:
        main()
    except BrokenPipeError:      # output piped into `head`
        pass
