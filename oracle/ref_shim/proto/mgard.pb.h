// TEST INFRASTRUCTURE. Plain-struct stand-in for the protoc-generated
// proto/mgard.pb.h (libprotobuf is not installed here).  It only has to satisfy
// the accessors the reference's MGARD-CPU headers name (reference
// src/mgard.proto:1-185 for the fields and enum values); nothing is serialised
// through it -- header bytes are checked against python protobuf instead
// (tests/test_cpu_convention.py).
#pragma once
#include <cstdint>
#include <istream>
#include <string>
#include <vector>

namespace google {
namespace protobuf {
using uint64 = std::uint64_t;
using uint8 = std::uint8_t;
template <typename T> class RepeatedField {
public:
  using const_iterator = typename std::vector<T>::const_iterator;
  void Resize(int n, const T &v) { items.resize(n, v); }
  T *mutable_data() { return items.data(); }
  int size() const { return (int)items.size(); }
  const_iterator begin() const { return items.begin(); }
  const_iterator end() const { return items.end(); }

private:
  std::vector<T> items;
};
} // namespace protobuf
} // namespace google

namespace mgard {
namespace pb {

#define SHIM_SCALAR(T, name)                                                   \
  T name##_v = T();                                                            \
  T name() const { return name##_v; }                                          \
  void set_##name(T v) { name##_v = v; }
#define SHIM_MESSAGE(T, name)                                                  \
  T name##_v;                                                                  \
  const T &name() const { return name##_v; }                                   \
  T *mutable_##name() { return &name##_v; }

struct VersionNumber {
  SHIM_SCALAR(std::uint64_t, major_)
  SHIM_SCALAR(std::uint64_t, minor_)
  SHIM_SCALAR(std::uint64_t, patch_)
};
struct CartesianGridTopology {
  SHIM_SCALAR(std::uint64_t, dimension)
  SHIM_MESSAGE(google::protobuf::RepeatedField<std::uint64_t>, shape)
};
struct ExplicitCubeGeometry {
  SHIM_MESSAGE(google::protobuf::RepeatedField<double>, coordinates)
  bool ParseFromIstream(std::istream *) { return false; }
};
struct Domain {
  enum Topology { CARTESIAN_GRID = 0 };
  enum Geometry { UNIT_CUBE = 0, EXPLICIT_CUBE = 1 };
  enum TopologyDefinitionCase {
    TOPOLOGY_DEFINITION_NOT_SET = 0,
    kCartesianGridTopology = 2
  };
  enum GeometryDefinitionCase {
    GEOMETRY_DEFINITION_NOT_SET = 0,
    kExplicitCubeGeometry = 4,
    kExplicitCubeFilename = 5
  };
  SHIM_SCALAR(Topology, topology)
  SHIM_SCALAR(Geometry, geometry)
  TopologyDefinitionCase topology_case = TOPOLOGY_DEFINITION_NOT_SET;
  GeometryDefinitionCase geometry_case = GEOMETRY_DEFINITION_NOT_SET;
  TopologyDefinitionCase topology_definition_case() const { return topology_case; }
  GeometryDefinitionCase geometry_definition_case() const { return geometry_case; }
  CartesianGridTopology grid;
  ExplicitCubeGeometry cube;
  std::string cube_filename;
  const CartesianGridTopology &cartesian_grid_topology() const { return grid; }
  CartesianGridTopology *mutable_cartesian_grid_topology() {
    topology_case = kCartesianGridTopology;
    return &grid;
  }
  const ExplicitCubeGeometry &explicit_cube_geometry() const { return cube; }
  ExplicitCubeGeometry *mutable_explicit_cube_geometry() {
    geometry_case = kExplicitCubeGeometry;
    return &cube;
  }
  const std::string &explicit_cube_filename() const { return cube_filename; }
};
struct Dataset {
  enum Type { FLOAT = 0, DOUBLE = 1 };
  SHIM_SCALAR(Type, type)
  SHIM_SCALAR(std::uint64_t, dimension)
};
struct ErrorControl {
  enum Mode { ABSOLUTE = 0, RELATIVE = 1 };
  enum Norm { L_INFINITY = 0, S_NORM = 1 };
  SHIM_SCALAR(Mode, mode)
  SHIM_SCALAR(Norm, norm)
  SHIM_SCALAR(double, s)
  SHIM_SCALAR(double, norm_of_original_data)
  SHIM_SCALAR(double, tolerance)
};
struct DomainDecomposition {
  enum Method { NOOP_METHOD = 0, MAX_DIMENSION = 1, BLOCK = 2, VARIABLE = 3 };
  SHIM_SCALAR(Method, method)
  SHIM_SCALAR(std::uint64_t, decomposition_dimension)
  SHIM_SCALAR(std::uint64_t, decomposition_size)
};
struct FunctionDecomposition {
  enum Transform { MULTILEVEL_COEFFICIENTS = 0 };
  enum Hierarchy {
    POWER_OF_TWO_PLUS_ONE = 0,
    MULTIDIMENSION_WITH_GHOST_NODES = 1,
    ONE_DIM_AT_A_TIME_WITH_GHOST_NODES = 2,
    HYBRID_HIERARCHY = 3
  };
  SHIM_SCALAR(Transform, transform)
  SHIM_SCALAR(Hierarchy, hierarchy)
  SHIM_SCALAR(std::uint64_t, l_target)
};
struct Quantization {
  enum Method { NOOP_QUANTIZATION = 0, COEFFICIENTWISE_LINEAR = 1 };
  enum BinWidths { PER_COEFFICIENT = 0, PER_LEVEL = 1 };
  enum Type { INT8_T = 0, INT16_T = 1, INT32_T = 2, INT64_T = 3 };
  SHIM_SCALAR(Method, method)
  SHIM_SCALAR(BinWidths, bin_widths)
  SHIM_SCALAR(Type, type)
  SHIM_SCALAR(bool, big_endian)
};
struct Encoding {
  enum Preprocessor { NOOP_PREPROCESSOR = 0, SHUFFLE = 1 };
  enum Compressor {
    NOOP_COMPRESSOR = 0,
    CPU_HUFFMAN_ZLIB = 1,
    CPU_HUFFMAN_ZSTD = 2,
    X_HUFFMAN = 3,
    X_HUFFMAN_LZ4 = 4,
    X_HUFFMAN_ZSTD = 5
  };
  SHIM_SCALAR(Preprocessor, preprocessor)
  SHIM_SCALAR(Compressor, compressor)
  SHIM_SCALAR(std::uint64_t, huffman_dictionary_size)
  SHIM_SCALAR(std::uint64_t, huffman_block_size)
};
struct Device {
  enum Backend { CPU = 0, X_SERIAL, X_OPENMP, X_CUDA, X_HIP, X_SYCL };
  SHIM_SCALAR(Backend, backend)
};
struct Header {
  SHIM_MESSAGE(VersionNumber, mgard_version)
  SHIM_MESSAGE(VersionNumber, file_format_version)
  SHIM_MESSAGE(Domain, domain)
  SHIM_MESSAGE(Dataset, dataset)
  SHIM_MESSAGE(ErrorControl, error_control)
  SHIM_MESSAGE(DomainDecomposition, domain_decomposition)
  SHIM_MESSAGE(FunctionDecomposition, function_decomposition)
  SHIM_MESSAGE(Quantization, quantization)
  SHIM_MESSAGE(Encoding, encoding)
  SHIM_MESSAGE(Device, device)
};
#undef SHIM_SCALAR
#undef SHIM_MESSAGE

} // namespace pb
} // namespace mgard
