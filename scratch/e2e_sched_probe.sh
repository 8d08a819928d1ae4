for c in 256 384 512; do for t in 0 32 64 128; do echo -n "chunk=$c tail=$t: "; ZIPGPU_CHUNK_ROWS=$c ZIPGPU_TAIL_ROWS=$t python scratch/e2e_zc.py | cut -c1-16; done; done
