for s in 0 1 2 4 3 6 7 8; do echo "SKIP=$s"; ARFE_FWD_SKIP=$s python bench.py --steps 6 --warmup 3 --no-cpu-baseline | python profiles/benchsum.py | grep roi_fuse_fwd; done
