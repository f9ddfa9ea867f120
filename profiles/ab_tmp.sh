python profiles/readloss_time.py 2>&1 | head -3
cd pinthememory_b200/_lib
nvcc -DPM_RL_MINB=3 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c ../csrc/pm_readloss9.cu -o pm_readloss9.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libpinmem_b200.so *.o
cd ../..
echo "--- 3 CTAs/SM"
python profiles/readloss_time.py 2>&1 | head -3
