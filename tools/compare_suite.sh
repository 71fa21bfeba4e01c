L=$1
export VOLPATH_B200_LIB=$PWD/$L
echo "== $L"
python tools/compare_ref_cuda.py --frames 64 --exact-bounds 2>/dev/null | tail -1
python tools/compare_ref_cuda.py --frames 64 --material 8 --exact-bounds 2>/dev/null | tail -1
python tools/compare_ref_cuda.py --frames 64 --material 4 --exact-bounds 2>/dev/null | tail -1
python tools/compare_ref_cuda.py --frames 64 --albedo 0.999 --density 3000 --exact-bounds 2>/dev/null | tail -1
python tools/compare_ref_cuda.py --frames 64 --julia --image 512 512 2>/dev/null | tail -1
