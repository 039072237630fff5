"""Dump the SASS of one kernel of libtilespmv_b200.so: python tools/sass_of.py <mangled-prefix> > out.sass"""
import re
import subprocess
import sys
txt = subprocess.run(["cuobjdump", "-sass", "tilespmv_b200/libtilespmv_b200.so"], capture_output=True, text=True).stdout
for p in re.split(r'\n\s*Function : ', txt):
    if p.startswith(sys.argv[1]):
        lines = [l for l in p.split('\n') if re.match(r'\s+/\*[0-9a-f]{4}\*/', l)]
        print('\n'.join(re.sub(r'/\* 0x[0-9a-f]+ \*/', '', l).strip() for l in lines))
        break
