"""Test infrastructure: CPU oracle of the MIMO U-Net hot path.  Never imported by the product."""
