#include "hdf5.h"
