"""Reference arm of bench.py: loader for the UNMODIFIED reference (a pristine copy under baseline/_ref/, made by
__graft_entry__.build() from /root/reference; git-ignored, shipped to the GPU box by gpurun).  Not product code."""
