"""Print registers / spills / smem per kernel from the build logs (rec_now_b200/csrc/_obj/*.o.log)."""
import glob, re, subprocess
for path in sorted(glob.glob("rec_now_b200/csrc/_obj/*.o.log")):
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:64]
        print(f"{path.split('/')[-1][:-6]:9s} {name:64s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill {m.group(3):>3s} smem {m.group(8)}")
