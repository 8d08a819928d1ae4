for ns in 0 300 600 1000 1500 2500; do echo -n "stagger16=$ns: "; ZIPGPU_STAGGER_NS=$ns python scratch/enc_only.py; done
