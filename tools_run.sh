timeout 600 python -m pytest tests/test_gpu_vae.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "csr or tokenizer" 2>&1 | tail -3
