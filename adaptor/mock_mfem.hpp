// mock_mfem.hpp -- the few dozen MFEM declarations heph_gpu_operator.hpp touches, backed by plain
// arrays, so that the adaptor can be compiled and driven end to end without MFEM (which cannot be
// built in this container).  Semantics follow MFEM 4.5/4.6 [MFEM-upstream]; this is TEST
// infrastructure, not a replacement for mfem.hpp.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#define MFEM_ABORT(msg) throw std::runtime_error(std::string("MFEM abort: ") + std::string(msg))

namespace mfem
{

template <class T>
class Array
{
public:
  Array() = default;
  explicit Array(int n) : _d(n) {}
  int Size() const { return (int)_d.size(); }
  void SetSize(int n) { _d.resize(n); }
  T & operator[](int i) { return _d[i]; }
  const T & operator[](int i) const { return _d[i]; }
  void Append(const T & v) { _d.push_back(v); }
  std::vector<T> _d;
};

class Vector
{
public:
  Vector() = default;
  explicit Vector(int n) : _d(n, 0.0) {}
  int Size() const { return (int)_d.size(); }
  void SetSize(int n) { _d.resize(n); }
  double & operator()(int i) { return _d[i]; }
  const double & operator()(int i) const { return _d[i]; }
  const double * HostRead() const { return _d.data(); }
  double * HostWrite() { return _d.data(); }
  double * HostReadWrite() { return _d.data(); }
  std::vector<double> _d;
};

class Operator
{
public:
  explicit Operator(int s = 0) : height(s), width(s) {}
  Operator(int h, int w) : height(h), width(w) {}
  virtual ~Operator() = default;
  int Height() const { return height; }
  int Width() const { return width; }
  virtual void Mult(const Vector & x, Vector & y) const = 0;
  /// mfem::Operator::ArrayMult: operator application on a set of vectors (default: one Mult each)
  virtual void ArrayMult(const std::vector<const Vector *> & X, std::vector<Vector *> & Y) const
  {
    for (size_t i = 0; i < X.size(); i++)
      Mult(*X[i], *Y[i]);
  }

protected:
  int height, width;
};

class Solver : public Operator
{
public:
  bool iterative_mode{false};
  virtual void SetOperator(const Operator & op) = 0;
};

struct IntegrationPoint
{
  double x{0}, y{0}, z{0}, weight{0};
  void Set3(double a, double b, double c) { x = a; y = b; z = c; }
};
class IntegrationRule
{
public:
  int GetNPoints() const { return (int)_p.size(); }
  const IntegrationPoint & IntPoint(int i) const { return _p[(size_t)i]; }
  std::vector<IntegrationPoint> _p;
};
/// MFEM's quadrature tables are not restated in the mock: per-quadrature-point sampling is compiled, not run, here
/// (the library side of it is exercised through the C ABI by tests/test_gpu_timestep.py)
class IntegrationRules
{
public:
  const IntegrationRule & Get(int, int) { MFEM_ABORT("mock MFEM has no integration rule tables"); }
};
static IntegrationRules IntRules;

struct Geometry
{
  enum Type { TETRAHEDRON = 4, CUBE = 5 };
  IntegrationPoint _c[6];
  Geometry()
  {
    _c[TETRAHEDRON].Set3(0.25, 0.25, 0.25);
    _c[CUBE].Set3(0.5, 0.5, 0.5);
  }
  const IntegrationPoint & GetCenter(int g) const { return _c[g]; }
};
static Geometry Geometries;

class ElementTransformation
{
public:
  int ElementNo{0}, Attribute{1};
  const IntegrationPoint * ip{nullptr};
  const std::vector<double> * xyz{nullptr};  // 8 or 4 corner nodes of the element
  int nv{8};
  void SetIntPoint(const IntegrationPoint * p) { ip = p; }
  int OrderW() const { return nv == 4 ? 0 : 2; }  // straight tets / trilinear hexes
  void Transform(const IntegrationPoint & p, Vector & x) const
  {
    x.SetSize(3);
    const double * X = xyz->data();
    if (nv == 4)
      for (int d = 0; d < 3; d++)
        x(d) = X[d] * (1 - p.x - p.y - p.z) + X[3 + d] * p.x + X[6 + d] * p.y + X[9 + d] * p.z;
    else
    {
      static const int HV[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
      for (int d = 0; d < 3; d++) x(d) = 0.0;
      for (int v = 0; v < 8; v++)
      {
        const double w = (HV[v][0] ? p.x : 1 - p.x) * (HV[v][1] ? p.y : 1 - p.y) * (HV[v][2] ? p.z : 1 - p.z);
        for (int d = 0; d < 3; d++) x(d) += w * X[3 * v + d];
      }
    }
  }
};

class Coefficient
{
public:
  virtual ~Coefficient() = default;
  virtual double Eval(ElementTransformation & T, const IntegrationPoint & ip) = 0;
};
/// piecewise constant by element attribute (mfem::PWConstCoefficient)
class PWConstCoefficient : public Coefficient
{
public:
  explicit PWConstCoefficient(std::vector<double> c) : _c(std::move(c)) {}
  double Eval(ElementTransformation & T, const IntegrationPoint &) override { return _c[T.Attribute - 1]; }
  std::vector<double> _c;
};

class FiniteElementCollection
{
public:
  explicit FiniteElementCollection(int p) : _p(p) {}
  int GetOrder() const { return _p; }
  int _p;
};

class FiniteElementSpace;
class GridFunction
{
public:
  FiniteElementSpace * FESpace() const { return nullptr; }
};
class FiniteElementSpace
{
public:
  int GetMaxElementOrder() const { return 1; }
};

class ParMesh
{
public:
  int kind{Geometry::CUBE};
  std::vector<double> verts;            // [nv][3]
  std::vector<int> elems;               // [ne][8|4]
  std::vector<int> attr;                // [ne]
  std::vector<int> elem_faces;          // [ne][4]  (tets)
  std::vector<int> face_verts;          // [nf][3]  (tets)
  int NV() const { return kind == Geometry::CUBE ? 8 : 4; }
  int GetNE() const { return (int)attr.size(); }
  int GetElementGeometry(int) const { return kind; }
  const GridFunction * GetNodes() const { return nullptr; }  // straight-sided mesh
  void GetElementVertices(int e, Array<int> & v) const
  {
    v.SetSize(NV());
    for (int k = 0; k < NV(); k++) v[k] = elems[(size_t)e * NV() + k];
  }
  const double * GetVertex(int i) const { return &verts[(size_t)i * 3]; }
  void GetElementFaces(int e, Array<int> & f, Array<int> & o) const
  {
    f.SetSize(4); o.SetSize(4);
    for (int k = 0; k < 4; k++) { f[k] = elem_faces[(size_t)e * 4 + k]; o[k] = 0; }
  }
  void GetFaceVertices(int f, Array<int> & v) const
  {
    v.SetSize(3);
    for (int k = 0; k < 3; k++) v[k] = face_verts[(size_t)f * 3 + k];
  }
  ElementTransformation * GetElementTransformation(int e)
  {
    _xyz.resize((size_t)NV() * 3);
    for (int k = 0; k < NV(); k++)
      for (int d = 0; d < 3; d++) _xyz[3 * k + d] = verts[(size_t)elems[(size_t)e * NV() + k] * 3 + d];
    _T.ElementNo = e; _T.Attribute = attr[e]; _T.xyz = &_xyz; _T.nv = NV();
    return &_T;
  }

private:
  ElementTransformation _T;
  std::vector<double> _xyz;
};

class Table
{
public:
  int RowSize(int) const { return 0; }
  const int * GetRow(int) const { return nullptr; }
};
class GroupCommunicator
{
public:
  const Table & GroupLDofTable() const { return _t; }
  Table _t;
};
class GroupTopology
{
public:
  int NGroups() const { return 1; }  // serial mock: only the local group
  int GetGroupMasterGroup(int) const { return 0; }
  int GetGroupMaster(int) const { return 0; }
  int GetNeighborRank(int) const { return 0; }
  int GetGroupSize(int) const { return 1; }
  const int * GetGroup(int) const { static int z = 0; return &z; }
};

// tensor-product elements expose the map lexicographic (tensor) index -> native local index (negative: -1-k, the basis
// function enters with a sign flip); an empty map means the native order is the tensor order
class FiniteElement
{
public:
  virtual ~FiniteElement() = default;
};
class TensorBasisElement
{
public:
  const Array<int> & GetDofMap() const { return _dof_map; }
  Array<int> _dof_map;
};
class MockTensorElement : public FiniteElement, public TensorBasisElement
{
};

class ParFiniteElementSpace
{
public:
  MockTensorElement fe;  // one element type per space
  const FiniteElement * GetFE(int) const { return &fe; }
  ParMesh * mesh{nullptr};
  FiniteElementCollection * fec{nullptr};
  int nd{0}, ndofs{0};
  std::vector<int> elem_dofs;  // [ne][nd] signed, MFEM local order
  ParMesh * GetParMesh() const { return mesh; }
  const FiniteElementCollection * FEColl() const { return fec; }
  void GetElementDofs(int e, Array<int> & d) const
  {
    d.SetSize(nd);
    for (int i = 0; i < nd; i++) d[i] = elem_dofs[(size_t)e * nd + i];
  }
  int GetVSize() const { return ndofs; }
  int GetTrueVSize() const { return ndofs; }
  int GetLocalTDofNumber(int ldof) const { return ldof; }
  int GetNRanks() const { return 1; }
  int GetMyRank() const { return 0; }
  GroupTopology & GetGroupTopo() { return _gt; }
  GroupCommunicator & GroupComm() { return _gc; }

private:
  GroupTopology _gt;
  GroupCommunicator _gc;
};

} // namespace mfem
