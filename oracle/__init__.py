"""TEST INFRASTRUCTURE ONLY -- see oracle/README.md. Nothing in libheif_b200/ may import this package."""
