// oracle/ref_stubs/istl_standins.hh -- TEST INFRASTRUCTURE ONLY (see README.md).
// Stand-ins for the dune-istl containers the reference's RowNormPreconditioner walks: a block CSR matrix whose row
// and column iterators know their index, and a block vector. Plus the grid-side names its typedefs mention.
#pragma once
#include <cstddef>
#include <vector>
#include <dune_standins.hh>

namespace Dune {

template <class B>
class BCRSMatrix {
 public:
  typedef std::size_t size_type;
  class ColIterator {
   public:
    ColIterator(std::vector<size_type>* c, std::vector<B>* b, size_type k) : c_(c), b_(b), k_(k) {}
    B& operator*() const { return (*b_)[k_]; }
    B* operator->() const { return &(*b_)[k_]; }
    size_type index() const { return (*c_)[k_]; }
    ColIterator& operator++() { ++k_; return *this; }
    bool operator!=(const ColIterator& o) const { return k_ != o.k_; }
   private:
    std::vector<size_type>* c_;
    std::vector<B>* b_;
    size_type k_;
  };
  struct Row {
    std::vector<size_type> cols;
    std::vector<B> blocks;
    B& operator[](size_type c) {
      for (size_type k = 0; k < cols.size(); ++k) if (cols[k] == c) return blocks[k];
      DUNE_THROW(Dune::RangeError, "no such block");
    }
    ColIterator begin() { return ColIterator(&cols, &blocks, 0); }
    ColIterator end() { return ColIterator(&cols, &blocks, cols.size()); }
  };
  class RowIterator {
   public:
    RowIterator(std::vector<Row>* r, size_type i) : r_(r), i_(i) {}
    Row& operator*() const { return (*r_)[i_]; }
    Row* operator->() const { return &(*r_)[i_]; }
    size_type index() const { return i_; }
    RowIterator& operator++() { ++i_; return *this; }
    bool operator!=(const RowIterator& o) const { return i_ != o.i_; }
   private:
    std::vector<Row>* r_;
    size_type i_;
  };
  // row-wise creation (only what Ax1Newton's optional matrix reordering names; never executed by the pins)
  class CreateIterator {
   public:
    CreateIterator(std::vector<Row>* r, size_type i) : r_(r), i_(i) {}
    size_type index() const { return i_; }
    void insert(size_type j) { (*r_)[i_].cols.push_back(j); (*r_)[i_].blocks.push_back(B()); }
    CreateIterator& operator++() { ++i_; return *this; }
    bool operator!=(const CreateIterator& o) const { return i_ != o.i_; }
   private:
    std::vector<Row>* r_;
    size_type i_;
  };
  typedef Row row_type;
  typedef B block_type;
  enum BuildMode { row_wise, random };
  BCRSMatrix() {}
  BCRSMatrix(size_type n, size_type, size_type, BuildMode) : rows(n) {}
  CreateIterator createbegin() { return CreateIterator(&rows, 0); }
  CreateIterator createend() { return CreateIterator(&rows, rows.size()); }
  size_type N() const { return rows.size(); }
  size_type nonzeroes() const { size_type k = 0; for (const Row& r : rows) k += r.cols.size(); return k; }
  Row& operator[](size_type i) { return rows[i]; }
  RowIterator begin() { return RowIterator(&rows, 0); }
  RowIterator end() { return RowIterator(&rows, rows.size()); }
  std::vector<Row> rows;
};

template <class B>
class BlockVector : public std::vector<B> {};

template <class GV, template <int> class Layout>
class MultipleCodimMultipleGeomTypeMapper {};

}  // namespace Dune
