# Build of the product library (CUDA, sm_100a) and of the test-only helpers.
#   make            -> ngspice-sf-mirror_b200/libngb200.so          (nvcc; the product)
#   make hostsim    -> tests/hostsim/libngb200_hostsim.so           (g++; CPU CI only)
#   make oracle     -> oracle/libngb_oracle.so + oracle/_ref/*      (gcc; checker only)
PKG := ngspice-sf-mirror_b200
CSRC := $(PKG)/csrc
NGB_INLINE_DIV ?= none
# divisions that share a denominator register share its reciprocal (tools/nvcc_outline.py): minimum number of
# divisions per denominator, 0 = off
NGB_SHARE_RCP ?= 0
NVCC ?= nvcc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -std=c++17 \
           -Xptxas -v -I$(CSRC) -Iinclude
HOSTC := $(CSRC)/ngb_host.c $(CSRC)/ngb_tran.c $(CSRC)/ngb_pivot.c $(CSRC)/ngb_b4temp.c
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: $(PKG)/libngb200.so

# the device compile goes through tools/nvcc_outline.py: plain nvcc steps, with the f64 div/rcp/sqrt
# of the BSIM4 kernel turned into calls of one shared subroutine each (instruction-fetch bound code)
$(PKG)/libngb200.so: $(CSRC)/ngb_cuda.cu $(HOSTC) $(HDRS) $(CSRC)/bsim4_finish.inc tools/nvcc_outline.py
	gcc -O2 -fPIC -fopenmp -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_host.c -o $(CSRC)/ngb_host.o
	gcc -O2 -fPIC -fopenmp -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_tran.c -o $(CSRC)/ngb_tran.o
	gcc -O2 -fPIC -fopenmp -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_pivot.c -o $(CSRC)/ngb_pivot.o
	gcc -O2 -fPIC -ffp-contract=off -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_b4temp.c -o $(CSRC)/ngb_b4temp.o
	python3 tools/nvcc_outline.py --outline-entries bsim4,ngb_k_b4_ --inline-div $(NGB_INLINE_DIV) --share-rcp $(NGB_SHARE_RCP) -- $(NVCC) $(NVFLAGS) $(NVDEFS) -c $(CSRC)/ngb_cuda.cu -o $(CSRC)/ngb_cuda.o 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; false)
	$(NVCC) -shared -o $@ $(CSRC)/ngb_cuda.o $(CSRC)/ngb_host.o $(CSRC)/ngb_tran.o $(CSRC)/ngb_pivot.o $(CSRC)/ngb_b4temp.o -lcudart -lgomp

hostsim: tests/hostsim/libngb200_hostsim.so
tests/hostsim/libngb200_hostsim.so: tests/hostsim/hostsim.cpp $(HOSTC) $(HDRS) $(CSRC)/bsim4_finish.inc
	gcc -O2 -fPIC -fopenmp -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_host.c -o tests/hostsim/ngb_host.o
	gcc -O2 -fPIC -fopenmp -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_tran.c -o tests/hostsim/ngb_tran.o
	gcc -O2 -fPIC -fopenmp -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_pivot.c -o tests/hostsim/ngb_pivot.o
	gcc -O2 -fPIC -ffp-contract=off -std=gnu99 -Wall -I$(CSRC) -Iinclude -x c -include $(CSRC)/c_compat.h -c $(CSRC)/ngb_b4temp.c -o tests/hostsim/ngb_b4temp.o
	g++ -O2 -fPIC -std=c++17 -ffp-contract=off -Wall -I$(CSRC) -Iinclude -x c++ -c tests/hostsim/hostsim.cpp -o tests/hostsim/hostsim.o
	g++ -shared -o $@ tests/hostsim/hostsim.o tests/hostsim/ngb_host.o tests/hostsim/ngb_tran.o tests/hostsim/ngb_pivot.o tests/hostsim/ngb_b4temp.o -lm -lgomp

clean:
	rm -f $(CSRC)/*.o $(PKG)/*.so tests/hostsim/*.o tests/hostsim/*.so
