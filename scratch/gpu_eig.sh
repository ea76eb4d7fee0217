echo "block deps"; timeout 120 python scratch/one_syevj.py 10; timeout 300 python scratch/time_big.py
echo "grid deps"; VVT_SYEVJ_GRIDDEP=1 timeout 120 python scratch/one_syevj.py 10; VVT_SYEVJ_GRIDDEP=1 timeout 300 python scratch/time_big.py
echo "block deps"; timeout 120 python scratch/one_syevj.py 10
echo "grid deps"; VVT_SYEVJ_GRIDDEP=1 timeout 120 python scratch/one_syevj.py 10
