// TEST INFRASTRUCTURE: stand-in for <boost/mpi/datatype.hpp> over the single-process mpi.h shim
#pragma once
#include <mpi.h>
namespace boost { namespace mpi {
template <class T> inline MPI_Datatype get_mpi_datatype(const T&) { return (MPI_Datatype)sizeof(T); }
} }
