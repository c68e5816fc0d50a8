cd $GRAFT_REPO_ROOT
for a in "fwd x 100 128" "fwd x 128 128" "fwd x 1024 784" "fwd x 512 256 200" "fwd x 256 512 37" "fwd p 1024 784" "fwd p 100 128" "fwd p 784 512" "bwd x 784 512" "bwd x 1024 784" "bwd p 784 512" "bwd p 1024 784" "bwd p 128 256" "bwd p 100 128"; do CGL_DEBUG_TMA=1 timeout 120 python profiles/tma_repro.py $a 2>&1 | tail -2 | cut -c1-160; done
T='python -m pytest tests/test_gpu_parity_report.py::test_md_parity_measured -x -q -k auto-mdgan_mnist'
timeout 300 $T 2>&1 | tail -1
