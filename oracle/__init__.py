"""oracle/ -- TEST INFRASTRUCTURE ONLY: CPU restatements of zinc's Zip commit path (C: zip_oracle.c,
Python: pyoracle.py).  The product (zinc_b200/) never imports this package."""
