// OcvYaml.hpp — reader for the subset of YAML 1.0 that cv::FileStorage emits, which is the on-disk format either side of
// the refinement path: detected_keypoints_data.yml (McCalib.cpp:493-548, read back by
// Calibration::loadDetectedKeypoints, :172-213) and the camera-parameter file read by
// Calibration::initializeCalibrationAllCam (:662-688). OpenCV itself is not available to the C++ side of this
// repository, so the format is parsed here: block mappings and block sequences by indentation, flow sequences
// "[ a, b, ... ]" and flow mappings "{ k: v }" that may wrap over several lines, quoted and plain scalars, and the
// "!!opencv-matrix" tag (a block mapping with rows / cols / dt / data). Anchors, multi-documents and block scalars are
// not part of what cv::FileStorage writes and are not supported.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace ocvyaml {

struct Node {
  enum Kind { NONE, SCALAR, SEQ, MAP } kind = NONE;
  std::string scalar;
  std::vector<Node> seq;
  std::vector<std::pair<std::string, Node>> map;   // insertion order kept
  std::string tag;

  const Node& operator[](const std::string& key) const {
    static const Node none;
    for (const auto& kv : map) if (kv.first == key) return kv.second;
    return none;
  }
  bool empty() const { return kind == NONE; }
  double asDouble(double def = 0.0) const { return kind == SCALAR ? std::strtod(scalar.c_str(), nullptr) : def; }
  int asInt(int def = 0) const { return kind == SCALAR ? (int)std::strtol(scalar.c_str(), nullptr, 10) : def; }
  // cv::FileNode >> std::vector<T>: a sequence of scalars
  std::vector<double> asDoubles() const { std::vector<double> v; for (const Node& n : seq) v.push_back(n.asDouble()); return v; }
  std::vector<int> asInts() const { std::vector<int> v; for (const Node& n : seq) v.push_back(n.asInt()); return v; }
  std::vector<std::string> asStrings() const { std::vector<std::string> v; for (const Node& n : seq) v.push_back(n.scalar); return v; }
  // cv::FileNode >> cv::Mat for "!!opencv-matrix": row-major data
  bool asMatrix(int& rows, int& cols, std::vector<double>& data) const {
    if (kind != MAP) return false;
    rows = (*this)["rows"].asInt(-1); cols = (*this)["cols"].asInt(-1);
    data = (*this)["data"].asDoubles();
    return rows >= 0 && cols >= 0 && (size_t)rows * (size_t)cols == data.size();
  }
};

class Parser {
 public:
  bool parseFile(const std::string& path, Node& root, std::string* err = nullptr) {
    std::ifstream f(path);
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::stringstream ss; ss << f.rdbuf();
    return parseText(ss.str(), root, err);
  }
  bool parseText(const std::string& text, Node& root, std::string* err = nullptr) {
    lines_.clear(); pos_ = 0; err_.clear();
    std::istringstream in(text);
    std::string l;
    while (std::getline(in, l)) {
      if (!l.empty() && l.back() == '\r') l.pop_back();
      if (l.rfind("%YAML", 0) == 0 || l == "---" || l == "...") continue;
      // comments: '#' at the start of the line or after white space, outside quotes (the reference's config files
      // carry one on almost every line)
      {
        bool q = false; char qc = 0;
        for (size_t i = 0; i < l.size(); ++i) {
          const char c = l[i];
          if (q) { if (c == '\\') ++i; else if (c == qc) q = false; continue; }
          if (c == '"' || c == '\'') { q = true; qc = c; continue; }
          if (c == '#' && (i == 0 || l[i - 1] == ' ' || l[i - 1] == '\t')) { l.erase(i); break; }
        }
        while (!l.empty() && (l.back() == ' ' || l.back() == '\t')) l.pop_back();
      }
      const size_t first = l.find_first_not_of(' ');
      if (first == std::string::npos) continue;
      lines_.push_back({(int)first, l.substr(first)});
    }
    root = Node();
    if (lines_.empty()) { root.kind = Node::MAP; return true; }
    root = parseBlock(lines_[0].indent);
    if (!err_.empty()) { if (err) *err = err_; return false; }
    return true;
  }

 private:
  struct Line { int indent; std::string text; };
  std::vector<Line> lines_;
  size_t pos_ = 0;
  std::string err_;

  static std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  static std::string unquote(const std::string& s) {
    if (s.size() >= 2 && s.front() == '"' && s.back() == '"') {
      std::string o;
      for (size_t i = 1; i + 1 < s.size(); ++i) {
        if (s[i] == '\\' && i + 2 < s.size()) { ++i; o += (s[i] == 'n' ? '\n' : s[i] == 't' ? '\t' : s[i]); }
        else o += s[i];
      }
      return o;
    }
    if (s.size() >= 2 && s.front() == '\'' && s.back() == '\'') return s.substr(1, s.size() - 2);
    return s;
  }
  // position of the ':' that ends a mapping key in a block line (outside quotes), or npos
  static size_t keyColon(const std::string& t) {
    bool q = false; char qc = 0;
    for (size_t i = 0; i < t.size(); ++i) {
      const char c = t[i];
      if (q) { if (c == '\\') ++i; else if (c == qc) q = false; continue; }
      if (c == '"' || c == '\'') { q = true; qc = c; continue; }
      if (c == '[' || c == '{') return std::string::npos;
      if (c == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
    }
    return std::string::npos;
  }
  // flow collection starting at `text` (first char '[' or '{'); may continue on the following lines
  Node parseFlowFrom(std::string text) {
    auto balanced = [](const std::string& s) {
      int d = 0; bool q = false; char qc = 0;
      for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) { if (c == '\\') ++i; else if (c == qc) q = false; continue; }
        if (c == '"' || c == '\'') { q = true; qc = c; }
        else if (c == '[' || c == '{') ++d;
        else if (c == ']' || c == '}') --d;
      }
      return d <= 0;
    };
    while (!balanced(text) && pos_ < lines_.size()) { text += " " + lines_[pos_].text; ++pos_; }
    size_t i = 0;
    return parseFlow(text, i);
  }
  Node parseFlow(const std::string& s, size_t& i) {
    auto skip = [&]() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i; };
    skip();
    Node n;
    if (i < s.size() && (s[i] == '[' || s[i] == '{')) {
      const bool is_map = s[i] == '{';
      const char close = is_map ? '}' : ']';
      n.kind = is_map ? Node::MAP : Node::SEQ;
      ++i;
      for (;;) {
        skip();
        if (i >= s.size()) { err_ = "unterminated flow collection"; return n; }
        if (s[i] == close) { ++i; break; }
        if (s[i] == ',') { ++i; continue; }
        if (is_map) {
          const size_t k0 = i;
          while (i < s.size() && s[i] != ':' && s[i] != close) ++i;
          const std::string key = unquote(trim(s.substr(k0, i - k0)));
          if (i < s.size() && s[i] == ':') ++i;
          n.map.emplace_back(key, parseFlow(s, i));
        } else {
          n.seq.push_back(parseFlow(s, i));
        }
      }
      return n;
    }
    // scalar up to , ] }
    const size_t b = i;
    if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
      const char qc = s[i]; ++i;
      while (i < s.size() && s[i] != qc) { if (s[i] == '\\') ++i; ++i; }
      if (i < s.size()) ++i;
    } else {
      while (i < s.size() && s[i] != ',' && s[i] != ']' && s[i] != '}') ++i;
    }
    n.kind = Node::SCALAR;
    n.scalar = unquote(trim(s.substr(b, i - b)));
    return n;
  }
  // value that follows "key:" or "- " on the current line (rest), or a nested block on the next lines
  Node parseValue(std::string rest, int parent_indent) {
    rest = trim(rest);
    std::string tag;
    if (rest.rfind("!!", 0) == 0) {
      const size_t sp = rest.find(' ');
      tag = rest.substr(0, sp);
      rest = sp == std::string::npos ? std::string() : trim(rest.substr(sp));
    }
    Node n;
    if (rest.empty()) {
      // cv::FileStorage puts an empty (or short) collection of a block parent on its own, deeper line: "key:\n   []"
      if (pos_ < lines_.size() && lines_[pos_].indent > parent_indent &&
          (lines_[pos_].text[0] == '[' || lines_[pos_].text[0] == '{')) {
        const std::string t = lines_[pos_].text;
        ++pos_;
        n = parseFlowFrom(t);
      } else if (pos_ < lines_.size() && (lines_[pos_].indent > parent_indent ||
                                   (lines_[pos_].indent == parent_indent && lines_[pos_].text.rfind("- ", 0) == 0)))
        n = parseBlock(lines_[pos_].indent);
      else { n.kind = Node::SCALAR; }
    } else if (rest[0] == '[' || rest[0] == '{') {
      n = parseFlowFrom(rest);
    } else {
      n.kind = Node::SCALAR; n.scalar = unquote(rest);
    }
    n.tag = tag;
    return n;
  }
  Node parseBlock(int indent) {
    Node n;
    if (pos_ >= lines_.size()) return n;
    const bool is_seq = lines_[pos_].text.rfind("- ", 0) == 0 || lines_[pos_].text == "-";
    n.kind = is_seq ? Node::SEQ : Node::MAP;
    while (pos_ < lines_.size() && lines_[pos_].indent == indent && err_.empty()) {
      const std::string t = lines_[pos_].text;
      if (is_seq) {
        if (!(t.rfind("- ", 0) == 0 || t == "-")) break;
        ++pos_;
        const std::string rest = t.size() > 2 ? trim(t.substr(2)) : std::string();
        const size_t kc = keyColon(rest);
        if (!rest.empty() && rest[0] != '[' && rest[0] != '{' && rest[0] != '"' && kc != std::string::npos) {
          // "- key: value" starts a mapping whose further keys are indented by two more columns
          Node m; m.kind = Node::MAP;
          m.map.emplace_back(unquote(trim(rest.substr(0, kc))), parseValue(rest.substr(kc + 1), indent + 2));
          while (pos_ < lines_.size() && lines_[pos_].indent == indent + 2) {
            const std::string t2 = lines_[pos_].text;
            const size_t k2 = keyColon(t2);
            if (k2 == std::string::npos) break;
            ++pos_;
            m.map.emplace_back(unquote(trim(t2.substr(0, k2))), parseValue(t2.substr(k2 + 1), indent + 2));
          }
          n.seq.push_back(m);
        } else {
          n.seq.push_back(parseValue(rest, indent));
        }
      } else {
        if (t.rfind("- ", 0) == 0) break;
        const size_t kc = keyColon(t);
        if (kc == std::string::npos) { err_ = "expected 'key: value' in: " + t; break; }
        ++pos_;
        n.map.emplace_back(unquote(trim(t.substr(0, kc))), parseValue(t.substr(kc + 1), indent));
      }
    }
    return n;
  }
};

}  // namespace ocvyaml
