for ns in 0 2000 4000 6000 8000 10000 14000 20000; do echo -n "stagger=$ns: "; ZIPGPU_STAGGER_NS=$ns python scratch/enc_only.py; done
