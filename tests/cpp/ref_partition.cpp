// The reference's own element-block partition of the ASM / Vanka smoother, printed: MeshASMPartitioning::DoPartition
// (MeshASMPartitioning.cpp:89-148) on every level of a refined mesh -- a generated HEX27 box or a Gambit file with several
// material groups (solid / porous / fluid classes) -- for a list of block sizes, together with every element's material
// flag in the reference's element order.  Compiled with the reference's own sources (oracle/ref_build); the JSON it
// prints is tests/golden/ref_partition.json (tests/golden/make_ref_stokes_golden.py).  Test infrastructure.
//
//   ref_partition box <nx> <ny> <nz> <levels> <block size> ...      |      ref_partition file <path.neu> <levels> <block size> ...
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#include "FemusInit.hpp"
#include "MultiLevelMesh.hpp"
#include "Mesh.hpp"
#include "MeshASMPartitioning.hpp"

using namespace femus;

int main(int argc, char** argv) {
  if (argc < 5) {
    std::cerr << "usage: " << argv[0] << " box <nx> <ny> <nz> <levels> <bs>... | file <path> <levels> <bs>...\n";
    return 1;
  }
  FemusInit init(argc, argv, MPI_COMM_WORLD);
  MultiLevelMesh ml_msh;
  int a = 2;
  const bool box = std::string(argv[1]) == "box";
  if (box) {
    ml_msh.GenerateCoarseBoxMesh(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]), 0., 1., 0., 1., 0., 1., HEX27, "seventh");
    a = 5;
  } else {
    ml_msh.ReadCoarseMesh(argv[2], "seventh", 1.);
    a = 3;
  }
  const unsigned nlevels = std::atoi(argv[a++]);
  ml_msh.RefineMesh(nlevels, nlevels, NULL);
  std::cout << "\nREF_PARTITION_JSON {\"levels\": [";
  for (unsigned l = 0; l < nlevels; l++) {
    Mesh* msh = ml_msh.GetLevel(l);
    const unsigned nel = msh->GetNumberOfElements();
    std::cout << (l ? ", " : "") << "{\"nel\": " << nel << ", \"material\": [";
    for (unsigned e = 0; e < nel; e++) std::cout << (e ? "," : "") << msh->GetElementMaterial(e);
    // one layer of near elements of every element (elem::BuildElementNearElement, Elem.cpp:493-526): what a Vanka block
    // takes its velocity dofs from
    std::cout << "], \"near\": [";
    for (unsigned e = 0; e < nel; e++) {
      std::cout << (e ? "," : "") << "[";
      const unsigned nn = msh->GetMeshElements()->GetElementNearElementSize(e, 1);
      for (unsigned j = 0; j < nn; j++) std::cout << (j ? "," : "") << msh->GetMeshElements()->GetElementNearElement(e, j);
      std::cout << "]";
    }
    std::cout << "], \"partitions\": [";
    for (int k = a; k < argc; k++) {
      const unsigned bs = std::atoi(argv[k]);
      const unsigned block_size[3] = {bs, bs, bs};
      std::vector<std::vector<unsigned> > blocks;
      std::vector<unsigned> range(3, 0);
      MeshASMPartitioning part(*msh);
      part.DoPartition(block_size, blocks, range);
      std::cout << (k > a ? ", " : "") << "{\"block_size\": " << bs << ", \"block_type_range\": [" << range[0] << "," << range[1] << "," << range[2]
                << "], \"blocks\": [";
      for (size_t b = 0; b < blocks.size(); b++) {
        std::cout << (b ? "," : "") << "[";
        for (size_t i = 0; i < blocks[b].size(); i++) std::cout << (i ? "," : "") << blocks[b][i];
        std::cout << "]";
      }
      std::cout << "]}";
    }
    std::cout << "]}";
  }
  std::cout << "]}" << std::endl;
  return 0;
}
