mkdir -p gpurun_out
for u in 31 63 95 127 159 191 255; do echo "U in [$((u-2)),$u] (pairs per lane $(( (u+32)/32 )))"; CTC_UMAX=$u timeout 120 python profiles/ctc_time.py 2>&1 | grep lattice; done | tee gpurun_out/c38_lattice_vs_width.txt
