set -x
for t in 8x16 4x32 2x64 1x128 16x8; do GSD_FORCE_TILE=$t python tools/microbench_conv.py 64 64 320 427 16 64; done
python tools/microbench_conv.py 64 64 320 427 16 64 nopool 1
python tools/microbench_conv.py 64 64 320 427 16 64 nopool 3
python tools/microbench_conv.py 128 128 160 213 16 128
python tools/microbench_conv.py 128 128 160 213 16 64
GSD_FORCE_TILE=1x128 python tools/microbench_conv.py 128 128 160 213 16 128
python tools/microbench_conv.py 256 256 80 106 16 256
python tools/microbench_conv.py 256 256 80 106 16 128
