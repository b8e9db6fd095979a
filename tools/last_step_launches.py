"""Cut the launches of ONE training step out of an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of an eager
bench run: a step contains exactly one flat AdamW launch (`adamw_flat_kernel`), so the step is everything after the
second-to-last AdamW launch up to and including the last one.
    python tools/last_step_launches.py all_launches.csv one_step.csv"""
import sys

lines = open(sys.argv[1]).read().splitlines()
hdr = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = lines[hdr + 1:]
marks = [i for i, l in enumerate(rows) if "adamw_flat_kernel" in l]
if len(marks) >= 2:
    step = rows[marks[-2] + 1:marks[-1] + 1]
else:                       # no optimizer launch in the list: fall back to the tail
    step = rows[-340:]
open(sys.argv[2], "w").write("\n".join([lines[hdr]] + step) + "\n")
print(f"{len(step)} launches in the last step", file=sys.stderr)
