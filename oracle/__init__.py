"""CPU oracle for the Dual-Space-NeRF render path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (`dual_space_nerf_b200`) may import this
package.  Allowed importers: tests/, __graft_entry__.smoke(), and the
cpu_baseline / `--impl reference` legs of bench.py.
"""
