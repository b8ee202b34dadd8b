for shape in "16 1280 1280" "4096 1280 1280" "16384 640 640" "65536 320 320" "1232 1280 768" "4096 10240 1280"; do
  python tools/one_gemm.py w4 $shape 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qdm_gemm --csv python tools/one_gemm.py w4 $shape 3 2>/dev/null | grep -E "qdm_gemm" | awk -F'","' -v s="$shape" '{print s, $5, $NF}' | tail -2
done
