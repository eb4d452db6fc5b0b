// mmf.hpp -- Matrix Market coordinate reader, API of the reference's
// include/io/mmf.hpp (MMF<I,V>, Elem, iterator, DoRead).
//
// Same observable behaviour (reference include/io/mmf.hpp:179-343 and
// src/mmf.cpp:6-44; restated in SURVEY.md appendix A.1), different machinery:
// the file is read in one piece and tokenised in place by a non-template
// scanner (cfs_spmv_b200/host/mmf.cpp); this header only converts the scanned
// triplets to the caller's index / value types. Entries are 1-based.
#ifndef MMF_HPP
#define MMF_HPP

#include <cstddef>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace cfs {
namespace io {

template <typename IndexType, typename ValueType> struct Elem {
  IndexType row;
  IndexType col;
  ValueType val;
};

// One line of `in`, trimmed of blanks and tabs at both ends and cut at single
// spaces. false at end of file; like the reference, a last line that is not
// terminated by '\n' counts as end of file.
bool DoRead(std::ifstream &in, std::vector<std::string> &arguments);

namespace detail {
// What the scanner hands back: the fully expanded, ordered entry list.
struct ScannedMatrix {
  long nr_rows, nr_cols, nr_declared;
  bool symmetric, col_wise, zero_based;
  std::vector<int> row, col; // 1-based
  std::vector<double> val;
};
// Reads and expands `filename`. Fatal problems print the reference's message
// on stdout and exit(1).
void scan_matrix_market(const std::string &filename, ScannedMatrix &out);
// Only the banner, comment and size lines of a file image (same checks, same
// fatal messages); entries_offset is where the entry lines start. The GPU
// ingest of CSRMatrix(filename) takes over from there.
struct MmfHeader {
  long nr_rows, nr_cols, nr_declared;
  bool symmetric, col_wise, zero_based;
  size_t entries_offset;
};
void scan_matrix_market_header(const char *image, size_t bytes,
                               MmfHeader &out);
} // namespace detail

template <typename IndexType, typename ValueType> class MMF {
public:
  typedef IndexType idx_t;
  typedef ValueType val_t;
  typedef Elem<IndexType, ValueType> elem_t;

  explicit MMF(const std::string &filename) {
    detail::ScannedMatrix s;
    detail::scan_matrix_market(filename, s);
    nr_rows_ = (IndexType)s.nr_rows;
    nr_cols_ = (IndexType)s.nr_cols;
    nr_declared_ = (IndexType)s.nr_declared;
    symmetric_ = s.symmetric;
    col_wise_ = s.col_wise;
    zero_based_ = s.zero_based;
    entries_.resize(s.row.size());
    for (size_t k = 0; k < entries_.size(); ++k) {
      entries_[k].row = (IndexType)s.row[k];
      entries_[k].col = (IndexType)s.col[k];
      entries_[k].val = (ValueType)s.val[k];
    }
  }

  IndexType GetNrRows() const { return nr_rows_; }
  IndexType GetNrCols() const { return nr_cols_; }
  // expanded count whenever the entries were materialised (always, unless the
  // header carries the `row` token of a general file)
  IndexType GetNrNonzeros() const {
    return (symmetric_ || col_wise_) ? (IndexType)entries_.size()
                                     : nr_declared_;
  }
  bool IsSymmetric() const { return symmetric_; }
  bool IsColWise() const { return col_wise_; }
  bool IsZeroBased() const { return zero_based_; }

  class iterator
      : public std::iterator<std::forward_iterator_tag, elem_t> {
  public:
    iterator() : owner_(0), at_(0) {}
    iterator(MMF *owner, size_t at) : owner_(owner), at_(at) {}
    bool operator==(const iterator &o) const {
      return owner_ == o.owner_ && at_ == o.at_;
    }
    bool operator!=(const iterator &o) const { return !(*this == o); }
    void operator++() { ++at_; }
    elem_t &operator*() { return owner_->entries_[at_]; }

  private:
    MMF *owner_;
    size_t at_;
  };

  iterator begin() { return iterator(this, 0); }
  iterator end() { return iterator(this, entries_.size()); }

private:
  IndexType nr_rows_, nr_cols_, nr_declared_;
  bool symmetric_, col_wise_, zero_based_;
  std::vector<elem_t> entries_;
};

} // namespace io
} // namespace cfs

#endif
