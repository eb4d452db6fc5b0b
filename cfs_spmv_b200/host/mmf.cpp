// mmf.cpp -- Matrix Market scanner behind include/io/mmf.hpp.
//
// Behavioural contract = the reference loader (include/io/mmf.hpp:179-343,
// src/mmf.cpp:6-44), checked against dumps of the compiled reference in
// tests/test_mmf_loader.py:
//   * a line is trimmed of ' ' and '\t' at both ends and split at single ' '
//     (empty tokens dropped); tabs INSIDE a line are not separators;
//   * a final line without '\n' is not seen;
//   * line 1: "%%MatrixMarket <obj> coordinate <field> general|symmetric
//     [base-0|base-1|column|row ...]"; <field> is ignored; a first token that
//     starts with "%%" but is not the banner is fatal; anything else means
//     "no banner": line 1 is already the size line (or a '%' comment);
//   * '%' lines directly before the size line are skipped;
//   * entries: >= 3 tokens -> atoi, atoi, atof; exactly 2 -> value 0.42;
//   * symmetric files are expanded (every off-diagonal entry mirrored,
//     duplicates kept), then everything is ordered by (row, col);
//   * fatal problems: message on stdout, exit(1).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>

#include "io/mmf.hpp"

namespace cfs {
namespace io {

namespace {

void die(const char *message) {
  std::cout << message << std::endl;
  exit(1);
}

// [begin, end) of one line inside the file image, without the '\n'
struct Line {
  const char *begin, *end;
};

class FileImage {
public:
  explicit FileImage(const std::string &path) : data_(0), size_(0), at_(0) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f)
      die("MMF file error.");
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    own_.resize(size > 0 ? (size_t)size : 0);
    if (size > 0 && fread(&own_[0], 1, (size_t)size, f) != (size_t)size)
      die("MMF file error.");
    fclose(f);
    data_ = own_.empty() ? 0 : &own_[0];
    size_ = own_.size();
  }
  // an image somebody else holds (the mapped file of the GPU ingest)
  FileImage(const char *image, size_t bytes)
      : data_(image), size_(bytes), at_(0) {}
  // next '\n'-terminated line; false when only an unterminated tail is left
  bool next(Line &l) {
    if (at_ >= size_)
      return false;
    const char *p = data_ + at_;
    const void *nl = memchr(p, '\n', size_ - at_);
    if (!nl)
      return false;
    l.begin = p;
    l.end = (const char *)nl;
    at_ = (size_t)(l.end - data_) + 1;
    return true;
  }
  int peek() const { return at_ < size_ ? data_[at_] : -1; }
  void skip_line() {
    Line l;
    if (!next(l))
      at_ = size_;
  }
  size_t position() const { return at_; }

private:
  const char *data_;
  size_t size_;
  std::vector<char> own_;
  size_t at_;
};

void trim(Line &l) {
  while (l.begin < l.end && (*l.begin == ' ' || *l.begin == '\t'))
    ++l.begin;
  while (l.end > l.begin && (l.end[-1] == ' ' || l.end[-1] == '\t'))
    --l.end;
}

// cuts a trimmed line at single spaces, dropping empty tokens
void tokens_of(Line l, std::vector<std::string> &out) {
  out.clear();
  trim(l);
  const char *p = l.begin;
  while (p < l.end) {
    const char *q = (const char *)memchr(p, ' ', (size_t)(l.end - p));
    if (!q)
      q = l.end;
    if (q > p)
      out.push_back(std::string(p, q));
    p = q + 1;
  }
}

// an entry line, without building std::strings: up to three tokens
inline int entry_of(Line l, long &r, long &c, double &v) {
  trim(l);
  const char *p = l.begin;
  const char *tok[3];
  int n = 0, total = 0;
  while (p < l.end) {
    const char *q = (const char *)memchr(p, ' ', (size_t)(l.end - p));
    if (!q)
      q = l.end;
    if (q > p) {
      if (n < 3)
        tok[n++] = p;
      ++total;
    }
    p = q + 1;
  }
  // the line is followed by '\n' in the image, so strtol/strtod stop there
  if (n >= 2) {
    r = (int)strtol(tok[0], 0, 10);
    c = (int)strtol(tok[1], 0, 10);
  }
  if (n >= 3)
    v = strtod(tok[2], 0);
  return total;
}

} // namespace

bool DoRead(std::ifstream &in, std::vector<std::string> &arguments) {
  std::string text;
  if (std::getline(in, text).eof())
    return false;
  Line l;
  l.begin = text.data();
  l.end = text.data() + text.size();
  tokens_of(l, arguments);
  return true;
}

namespace detail {

namespace {

// banner, comments and the size line; leaves `file` at the first entry line
void scan_header(FileImage &file, ScannedMatrix &m) {
  m.symmetric = false;
  m.col_wise = true;
  m.zero_based = false;
  std::vector<std::string> args;
  Line line;
  bool banner_less = false;

  // ---- line 1
  bool have = file.next(line);
  if (have)
    tokens_of(line, args);
  if (!have || args.empty())
    die("size line error in MMF file.");
  if (args[0] != "%%MatrixMarket") {
    if (args[0].size() > 2 && args[0][0] == '%' && args[0][1] == '%')
      die("invalid header line in MMF file.");
    banner_less = true;
  } else {
    if (args.size() < 5)
      die("less arguments in header line of MMF file.");
    if (args[2] != "coordinate")
      die("unsupported matrix format in header line of MMF file.");
    if (args[4] == "general")
      m.symmetric = false;
    else if (args[4] == "symmetric")
      m.symmetric = true;
    else
      die("unsupported symmetry in header line of MMF file.");
    for (size_t i = 5; i < args.size(); ++i) {
      if (args[i] == "base-0")
        m.zero_based = true;
      else if (args[i] == "base-1")
        m.zero_based = false;
      else if (args[i] == "column")
        m.col_wise = true;
      else if (args[i] == "row")
        m.col_wise = false;
    }
  }

  // ---- size line
  if (!banner_less || args[0][0] == '%') {
    while (file.peek() == '%')
      file.skip_line();
    if (!file.next(line))
      die("size line error in MMF file.");
    tokens_of(line, args);
  }
  if (args.size() < 2) {
    if (!args.empty())
      std::cout << args[0] << std::endl;
    die("bad input, less arguments in line of MMF file.");
  }
  m.nr_rows = atoi(args[0].c_str());
  m.nr_cols = atoi(args[1].c_str());
  // the reference reads the count through its entry parser: a value, 0.42 when
  // the third token is missing, truncated to the index type
  m.nr_declared = (long)(args.size() >= 3 ? atof(args[2].c_str()) : 0.42);
}

} // namespace

void scan_matrix_market_header(const char *image, size_t bytes,
                               MmfHeader &out) {
  FileImage file(image, bytes);
  ScannedMatrix m;
  scan_header(file, m);
  out.nr_rows = m.nr_rows;
  out.nr_cols = m.nr_cols;
  out.nr_declared = m.nr_declared;
  out.symmetric = m.symmetric;
  out.col_wise = m.col_wise;
  out.zero_based = m.zero_based;
  out.entries_offset = file.position();
}

void scan_matrix_market(const std::string &filename, ScannedMatrix &m) {
  FileImage file(filename);
  scan_header(file, m);
  std::vector<std::string> args;
  Line line;
  const long declared = m.nr_declared;

  // ---- entries
  const size_t reserve = (size_t)(declared > 0 ? declared : 0);
  std::vector<int> row, col;
  std::vector<double> val;
  row.reserve(m.symmetric ? 2 * reserve : reserve);
  col.reserve(m.symmetric ? 2 * reserve : reserve);
  val.reserve(m.symmetric ? 2 * reserve : reserve);
  for (long k = 0; k < declared; ++k) {
    if (!file.next(line))
      die(m.symmetric || m.col_wise ? "Requesting dereference, but mmf ended."
                                    : "Requesting dereference, but mmf ended");
    long r = 0, c = 0;
    double v = 0.42; // value of two-token ("pattern") lines
    const int ntok = entry_of(line, r, c, v);
    if (ntok < 2) {
      tokens_of(line, args);
      if (!args.empty())
        std::cout << args[0] << std::endl;
      die("bad input, less arguments in line of MMF file.");
    }
    if (m.zero_based) {
      ++r;
      ++c;
    }
    row.push_back((int)r);
    col.push_back((int)c);
    val.push_back(v);
    if (m.symmetric && r != c) {
      row.push_back((int)c);
      col.push_back((int)r);
      val.push_back(v);
    }
  }

  // ---- order by (row, col); a general file announced as `row` ordered is
  // taken in file order, as the reference streams it
  if (m.symmetric || m.col_wise) {
    std::vector<size_t> order(row.size());
    std::iota(order.begin(), order.end(), (size_t)0);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      if (row[a] != row[b])
        return row[a] < row[b];
      return col[a] < col[b];
    });
    m.row.resize(order.size());
    m.col.resize(order.size());
    m.val.resize(order.size());
    for (size_t k = 0; k < order.size(); ++k) {
      m.row[k] = row[order[k]];
      m.col[k] = col[order[k]];
      m.val[k] = val[order[k]];
    }
  } else {
    m.row.swap(row);
    m.col.swap(col);
    m.val.swap(val);
  }
}

} // namespace detail

} // namespace io
} // namespace cfs
