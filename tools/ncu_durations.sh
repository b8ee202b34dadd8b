# pure GPU durations of the GEMM kernels for a few shapes (ncu, cold cache, serialised)
for shape in "$@"; do
  python tools/one_gemm.py w4 $shape 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qdm_gemm|smallm" --csv python tools/one_gemm.py w4 $shape 3 2>/dev/null | grep -E "qdm_gemm|smallm" | awk -F'","' -v s="$shape" '{print s, substr($5,1,60), $NF}' | tail -1
done
