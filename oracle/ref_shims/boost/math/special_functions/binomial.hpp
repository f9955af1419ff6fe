// TEST INFRASTRUCTURE: empty stand-in (uq.cpp is not part of the oracle build)
