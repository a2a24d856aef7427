"""apoget_b200 -- B200-native species-assignment path of APOGeT.

The product is ``libapt_b200.so`` (CUDA kernels + ordered resolution engine behind the C ABI of
``include/apt_b200.h``) and the C++ drop-in ``include/apt/phylogenetictree.hpp``.  This package is the
Python harness around the C ABI used by the tests and the benchmark:

* ``apoget_b200.tree.PhylogeneticTree`` -- the reference's tree interface for device-encoded genomes;
* ``apoget_b200.kernels``               -- the stateless kernels (pairs scoring, dense sweep, criterion, pairwise);
* ``apoget_b200.synth``                 -- synthetic host simulations (birth traces);
* ``apoget_b200.build``                 -- nvcc build of the extension (sm_100a).
"""
__all__ = ["synth", "tree", "kernels", "build"]
