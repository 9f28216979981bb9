// host/spm_encode_main.cc -- the `spm_encode` caller re-looped for batches.
//
// The reference CLI (src/spm_encode_main.cc:102-165) encodes one line per call in a
// single-threaded getline loop.  This drop-in keeps its flags for the accelerated
// formats (--model, --input, --output, --output_format=id|piece, --extra_options) but
// reads --batch_lines lines at a time into one packed buffer, makes ONE batch call into
// the engine, and writes the lines back in order.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "host/sentencepiece_processor.h"

namespace {
bool Flag(const char *arg, const char *name, std::string *val) {
  const size_t l = strlen(name);
  if (strncmp(arg, name, l) == 0 && arg[l] == '=') { *val = arg + l + 1; return true; }
  return false;
}
}  // namespace

int main(int argc, char **argv) {
  std::string model, input, output, format = "piece", extra, batch = "1000000", device = "0";
  bool verbose = false;
  for (int i = 1; i < argc; ++i) {
    std::string v;
    if (Flag(argv[i], "--model", &model) || Flag(argv[i], "--input", &input) || Flag(argv[i], "--output", &output) ||
        Flag(argv[i], "--output_format", &format) || Flag(argv[i], "--extra_options", &extra) ||
        Flag(argv[i], "--batch_lines", &batch) || Flag(argv[i], "--device", &device))
      continue;
    if (!strcmp(argv[i], "--verbose")) { verbose = true; continue; }
    if (argv[i][0] != '-') { input = argv[i]; continue; }
    std::cerr << "unknown flag " << argv[i] << "\n";
    return 2;
  }
  if (model.empty()) { std::cerr << "--model is required\n"; return 2; }
  if (format != "id" && format != "piece") {
    std::cerr << "--output_format=" << format << " is not on the accelerated path (id | piece)\n";
    return 2;
  }
  sentencepiece::SentencePieceProcessor sp;
  sp.SetDevice(std::stoi(device));
  auto st = sp.Load(model);
  if (!st.ok()) { std::cerr << st.ToString() << "\n"; return 1; }
  st = sp.SetEncodeExtraOptions(extra);
  if (!st.ok()) { std::cerr << st.ToString() << "\n"; return 1; }

  std::ifstream fin;
  if (!input.empty()) { fin.open(input, std::ios::binary); if (!fin) { std::cerr << "cannot open " << input << "\n"; return 1; } }
  std::istream &in = input.empty() ? std::cin : fin;
  std::ofstream fout;
  if (!output.empty()) { fout.open(output, std::ios::binary); if (!fout) { std::cerr << "cannot open " << output << "\n"; return 1; } }
  std::ostream &out = output.empty() ? std::cout : fout;

  const size_t batch_lines = std::max<size_t>(1, std::stoull(batch));
  std::vector<std::string> lines;
  std::vector<std::string_view> views;
  std::string line, buf;
  double enc_s = 0;
  size_t total = 0;
  bool eof = false;
  while (!eof) {
    lines.clear();
    while (lines.size() < batch_lines) {
      if (!std::getline(in, line)) { eof = true; break; }
      lines.push_back(line);
    }
    if (lines.empty()) break;
    views.assign(lines.begin(), lines.end());
    const auto t0 = std::chrono::steady_clock::now();
    buf.clear();
    if (format == "id") {
      std::vector<std::vector<int>> ids;
      st = sp.Encode(views, &ids);
      if (!st.ok()) { std::cerr << st.ToString() << "\n"; return 1; }
      for (const auto &v : ids) {
        for (size_t k = 0; k < v.size(); ++k) { if (k) buf += ' '; buf += std::to_string(v[k]); }
        buf += '\n';
      }
    } else {
      std::vector<std::vector<std::string>> pcs;
      st = sp.Encode(views, &pcs);
      if (!st.ok()) { std::cerr << st.ToString() << "\n"; return 1; }
      for (const auto &v : pcs) {
        for (size_t k = 0; k < v.size(); ++k) { if (k) buf += ' '; buf += v[k]; }
        buf += '\n';
      }
    }
    enc_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out.write(buf.data(), buf.size());
    total += lines.size();
  }
  if (verbose) fprintf(stderr, "encoded %zu lines, %.3f s in Encode+format\n", total, enc_s);
  return 0;
}
