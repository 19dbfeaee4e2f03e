#!/usr/bin/env python3
"""Registers / spills of every eri_class_kernel instantiation from nvcc -Xptxas -v logs.  usage: tools/ptxas_table.py build/obj/*.ptxas.log [filter]"""
import re, sys
logs = [a for a in sys.argv[1:] if a.endswith('.log')]
flt = [a for a in sys.argv[1:] if not a.endswith('.log')]
for path in logs:
    name = None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1); spill = None; continue
        m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
        if m: spill = m.groups(); continue
        m = re.search(r'Used (\d+) registers', line)
        if m and name:
            k = re.search(r'eri_class_kernelI((?:Li\d+E)+)NS_(\d+)(\w+?)E', name)
            if k:
                nums = re.findall(r'Li(\d+)E', k.group(1)); sink = k.group(3)[:int(k.group(2))]
                tag = '(%s%s|%s%s) fixr%s cb%s+%s kc%s+%s %s' % ('spd'[int(nums[0])], 'spd'[int(nums[1])], 'spd'[int(nums[2])], 'spd'[int(nums[3])], *nums[4:9], sink)
                if not flt or all(f in tag for f in flt):
                    print('%-52s regs %3s  stack %4s spill st/ld %4s/%4s' % (tag, m.group(1), spill[0], spill[1], spill[2]))
            name = None
