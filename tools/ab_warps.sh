# more, lighter warps for short block rows? 80-register kernel with up to 24 warps and smaller chunks vs the default shape
for w in "lap2d --grid 1024" "lap3d27 --grid 96" "banded --n 1048576" "band_contig --n 1048576"; do
  python tools/ab.py --workload $w --configs "2:0,2:21:4096,2:24:3072,2:24:2560,2:22:3584" --rounds 2 --iters 300
done
