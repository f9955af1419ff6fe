/* TEST INFRASTRUCTURE: empty stand-in (oracle/ref_shims/FemusConfig.hpp carries the opaque PETSc types) */
