"""Summarise `nvcc -Xptxas -v` output: registers / spills per kernel (filter by substring)."""
import re, subprocess, sys
txt = open(sys.argv[1]).read(); filt = sys.argv[2] if len(sys.argv) > 2 else ''
ent = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
dem = subprocess.run(['c++filt'] + [e[0] for e in ent], capture_output=True, text=True).stdout.split('\n')
for e, d in zip(ent, dem):
    if filt in d:
        d = d.replace('gsb::', '').replace('(int)', '').replace('(bool)', '').replace('(unsigned int)', '')
        print(f"regs {e[4]:>3} stack {e[1]:>4} spill {e[2]:>4}/{e[3]:>4}  {d[:100]}")
