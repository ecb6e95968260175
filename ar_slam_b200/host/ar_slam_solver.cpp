// ar_slam_solver.cpp -- host data store, schedules and map.yaml of the drop-in
// facade (see ar_slam_solver.hpp).  Follows the behaviour of the reference's
// ar_slam/src/ar_slam_util.cpp: loadYaml :304-368, saveYaml :371-465,
// addDetections :591-627, solveIncremental :629-678, solveCapture :680-742,
// solve :744-866, addConnectedCaptures :869-885, localizeMany/One :888-979,
// optimize :1001-1018, resetProblem :1021-1025 -- with the Ceres calls
// replaced by the C-ABI of the CUDA library.
#include "ar_slam_solver.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "../csrc/model.cuh"  // host build of the seed heuristics (calcInitValues & co.)

namespace {

// ---- a reader for the subset of YAML that saveYaml / yaml-cpp emit -----------
std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string unquote(std::string s) {
  s = trim(s);
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}
std::vector<double> parse_flow_seq(const std::string& v) {
  const size_t a = v.find('['), b = v.rfind(']');
  if (a == std::string::npos || b == std::string::npos || b < a) throw std::runtime_error("expected a flow sequence: " + v);
  std::vector<double> out;
  std::stringstream ss(v.substr(a + 1, b - a - 1));
  std::string item;
  while (std::getline(ss, item, ',')) {
    item = trim(item);
    if (item.empty()) continue;
    if (item == ".nan" || item == ".NaN") out.push_back(std::nan(""));
    else if (item == ".inf") out.push_back(INFINITY);
    else if (item == "-.inf") out.push_back(-INFINITY);
    else out.push_back(std::strtod(item.c_str(), nullptr));
  }
  return out;
}
struct YLine { int indent; bool item; std::string key, value; };
std::vector<YLine> read_lines(const std::string& fn) {
  std::ifstream f(fn);
  if (!f) throw std::runtime_error("cannot open " + fn);
  std::vector<YLine> out;
  std::string raw;
  while (std::getline(f, raw)) {
    std::string s = raw;
    if (trim(s).empty() || trim(s)[0] == '#' || trim(s) == "---") continue;
    YLine l;
    l.indent = (int)s.find_first_not_of(' ');
    std::string body = trim(s);
    l.item = body.rfind("- ", 0) == 0;
    if (l.item) { body = trim(body.substr(2)); l.indent += 2; }
    const size_t c = body.find(':');
    if (c == std::string::npos) throw std::runtime_error("cannot parse yaml line: " + raw);
    l.key = unquote(body.substr(0, c));
    l.value = trim(body.substr(c + 1));
    out.push_back(l);
  }
  return out;
}

// yaml-cpp writes doubles with max_digits10 significant digits (lossless)
std::string num(double v) {
  if (std::isnan(v)) return ".nan";
  if (std::isinf(v)) return v > 0 ? ".inf" : "-.inf";
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.17g", v);
  return buf;
}

void check(arslam_solver* s, int rc, const char* what) {
  if (rc != ARSLAM_OK) throw std::runtime_error(std::string(what) + ": " + arslam_last_error(s));
}

}  // namespace

bool endswith(const std::string& str, const std::string& suffix) {
  return str.size() >= suffix.size() && str.compare(str.size() - suffix.size(), suffix.size(), suffix) == 0;
}
std::string filename_no_ext(std::string filepath) {
  const size_t slash = filepath.find_last_of('/');
  if (slash != std::string::npos) filepath = filepath.substr(slash + 1);
  const size_t dot = filepath.find_last_of('.');
  if (dot != std::string::npos) filepath = filepath.substr(0, dot);
  return filepath;
}

ArSlamSolver::ArSlamSolver() { arslam_default_options(&options_); }
ArSlamSolver::~ArSlamSolver() { if (gpu_) arslam_destroy(gpu_); }

arslam_solver* ArSlamSolver::handle() {
  if (!gpu_) {
    const char* dev = std::getenv("ARSLAM_DEVICE");
    const int rc = arslam_create(dev ? std::atoi(dev) : 0, &options_, &gpu_);
    if (rc != ARSLAM_OK) throw std::runtime_error(std::string("arslam_create: ") + arslam_last_error(nullptr));
  }
  return gpu_;
}

Capture& ArSlamSolver::addCapture(CaptureUid cap_uid, std::string fn) {
  CaptureHandle h(captures_.size());
  if (!capture_map_.try_emplace(cap_uid, h).second) throw std::runtime_error("Capture with uid already added");
  captures_.emplace_back(std::move(cap_uid), h, std::move(fn));
  return captures_.back();
}
Aruco& ArSlamSolver::addAruco(ArucoId ar_id) {
  ArucoHandle h(arucos_.size());
  aruco_map_.try_emplace(ar_id, h);  // an existing id keeps its first handle, like the reference
  arucos_.emplace_back(std::move(ar_id), h);
  return arucos_.back();
}
Aruco& ArSlamSolver::getOrAddAruco(const ArucoId& ar_id) {
  auto it = aruco_map_.find(ar_id);
  return it == aruco_map_.end() ? addAruco(ar_id) : at(it->second);
}
Block& ArSlamSolver::addBlock(const ArucoRect& rect, CaptureHandle c, ArucoHandle a) {
  BlockHandle h(blocks_.size());
  blocks_.emplace_back(h, rect, c, a);
  at(c).blocks.emplace_back(h);
  at(a).blocks.emplace_back(h);
  return blocks_.back();
}

// ---- image ingest (reference ar_slam_util.cpp:218-286) ----------------------------------------------------
namespace {
struct Netpbm { int width = 0, height = 0, channels = 0; std::vector<uint8_t> data; };   // data: BGR or grey
Netpbm read_netpbm(const std::string& fn) {
  std::ifstream in(fn, std::ios::binary);
  if (!in) throw std::runtime_error("error loading image " + fn);
  std::string magic;
  in >> magic;
  if (magic != "P5" && magic != "P6") throw std::runtime_error("error loading image " + fn + " (binary netpbm P5 / P6 only)");
  int vals[3], got = 0;
  while (got < 3) {
    int c = in.peek();
    if (c == '#') { std::string skip; std::getline(in, skip); continue; }
    if (std::isspace(c)) { in.get(); continue; }
    if (!(in >> vals[got++])) throw std::runtime_error("error loading image " + fn);
  }
  in.get();                                     // the single whitespace before the raster
  if (vals[2] != 255 || vals[0] < 1 || vals[1] < 1) throw std::runtime_error("error loading image " + fn + " (8-bit only)");
  Netpbm img;
  img.width = vals[0]; img.height = vals[1]; img.channels = magic == "P6" ? 3 : 1;
  img.data.resize((size_t)img.width * img.height * img.channels);
  in.read(reinterpret_cast<char*>(img.data.data()), (std::streamsize)img.data.size());
  if ((size_t)in.gcount() != img.data.size()) throw std::runtime_error("error loading image " + fn + " (truncated)");
  if (img.channels == 3)                        // netpbm is RGB, the detector takes cv::imread's BGR
    for (size_t i = 0; i < img.data.size(); i += 3) std::swap(img.data[i], img.data[i + 2]);
  return img;
}
Netpbm rotate_90_clockwise(const Netpbm& a) {   // cv::rotate(ROTATE_90_CLOCKWISE)
  Netpbm r;
  r.width = a.height; r.height = a.width; r.channels = a.channels;
  r.data.resize(a.data.size());
  for (int y = 0; y < a.height; ++y)
    for (int x = 0; x < a.width; ++x)
      for (int c = 0; c < a.channels; ++c)
        r.data[((size_t)x * r.width + (a.height - 1 - y)) * a.channels + c] = a.data[((size_t)y * a.width + x) * a.channels + c];
  return r;
}
}  // namespace

void ArSlamSolver::loadImages(const std::vector<std::string>& img_fns) {
  arslam_detect_params params;
  arslam_detect_default_params(&params);
  params.min_corner_distance_rate = 0.1;        // :250
  arslam_detector* det = nullptr;
  for (const auto& img_fn : img_fns) {
    Netpbm img = read_netpbm(img_fn);
    // checkAndFixImageSize (:218-245)
    if (camera_.size.has_value()) {
      if (img.width == camera_.size->second && img.height == camera_.size->first && img.width != img.height) {
        std::cerr << "WARNING : some images are rotated relative to others fixing by rotating 90 degrees" << std::endl;
        img = rotate_90_clockwise(img);
      }
      if (std::make_pair(img.width, img.height) != camera_.size.value()) {
        if (det) arslam_detector_destroy(det);
        std::ostringstream ss;
        ss << "Loaded images should all be same size :  expected [" << camera_.size->first << " x " << camera_.size->second
           << "] got [" << img.width << " x " << img.height << "]";
        throw std::runtime_error(ss.str());
      }
    } else {
      camera_.size = std::make_pair(img.width, img.height);
    }
    if (!det && arslam_detector_create(0, 1, camera_.size->first, camera_.size->second, &det) != ARSLAM_OK)
      throw std::runtime_error(std::string("arslam_detector_create: ") + arslam_detector_last_error(nullptr));
    const int cap = 1024;
    int32_t n = 0;
    std::vector<int32_t> ids(cap);
    std::vector<float> corners((size_t)cap * 8);
    if (arslam_detect_markers(det, img.data.data(), 1, img.width, img.height, img.channels, 0, &params, cap, &n, ids.data(),
                              corners.data()) != ARSLAM_OK) {
      std::string msg = std::string("arslam_detect_markers: ") + arslam_detector_last_error(det);
      arslam_detector_destroy(det);
      throw std::runtime_error(msg);
    }
    if (n <= 2) std::cout << "Warning not enough AR tags detected in " << img_fn << std::endl;   // :270-272
    Capture& capture = addCapture(genUniqueCaptureUid(), img_fn);
    for (int k = 0; k < n; ++k) {
      Aruco& aruco = getOrAddAruco(ArucoId("aruco_4X4_50_" + std::to_string(ids[k])));
      ArucoRect rect;                           // ArucoRect(rects, img.size()): from_cv_img in double (hpp:257-282)
      for (int j = 0; j < 4; ++j) {
        rect.corners[j].x = corners[(size_t)k * 8 + 2 * j] - 0.5 * img.width;
        rect.corners[j].y = corners[(size_t)k * 8 + 2 * j + 1] - 0.5 * img.height;
      }
      addBlock(rect, capture.handle, aruco.handle);
    }
  }
  if (det) arslam_detector_destroy(det);
}

CaptureUid ArSlamSolver::genUniqueCaptureUid() const {
  const std::string base = "cap_" + std::to_string(captures_.size());
  if (CaptureUid uid{base}; capture_map_.count(uid) == 0) return uid;
  for (unsigned i = 0; i < 1000; ++i) {
    CaptureUid uid{base + "_" + std::to_string(i)};
    if (capture_map_.count(uid) == 0) return uid;
  }
  throw std::runtime_error("cannot generate unique id");
}

void ArSlamSolver::loadYaml(const std::string& fn) {
  const std::vector<YLine> lines = read_lines(fn);
  struct Blk { std::string cap, tag; std::vector<double> rect; bool has_rect = false; };
  std::vector<Blk> blks;
  std::vector<std::pair<std::string, std::pair<std::vector<double>, std::string>>> caps;
  std::vector<std::pair<std::string, std::vector<double>>> tags;
  std::vector<double> cam_params;
  int width = -1, height = -1;
  std::string section, entry;
  for (const YLine& l : lines) {
    if (l.indent == 0) { section = l.key; entry.clear(); continue; }
    if (section == "blocks") {
      if (l.item) blks.emplace_back();
      if (blks.empty()) throw std::runtime_error("malformed blocks section in " + fn);
      Blk& b = blks.back();
      if (l.key == "capture") b.cap = unquote(l.value);
      else if (l.key == "aruco") b.tag = unquote(l.value);
      else if (l.key == "aruco_rect") { b.rect = parse_flow_seq(l.value); b.has_rect = true; }
    } else if (section == "captures") {
      if (l.indent == 2) { caps.push_back({l.key, {{}, ""}}); continue; }
      if (caps.empty()) throw std::runtime_error("malformed captures section in " + fn);
      if (l.key == "inv_pose") caps.back().second.first = parse_flow_seq(l.value);
      else if (l.key == "img_fn") caps.back().second.second = unquote(l.value);
    } else if (section == "arucos") {
      if (l.indent == 2) { tags.push_back({l.key, {}}); continue; }
      if (tags.empty()) throw std::runtime_error("malformed arucos section in " + fn);
      if (l.key == "pose") tags.back().second = parse_flow_seq(l.value);
    } else if (section == "camera") {
      if (l.key == "params") cam_params = parse_flow_seq(l.value);
      else if (l.key == "width") width = std::atoi(l.value.c_str());
      else if (l.key == "height") height = std::atoi(l.value.c_str());
    }
  }
  // same order as the reference: captures, arucos, blocks, camera
  for (auto& c : caps) {
    CaptureUid uid(c.first);
    if (capture_map_.count(uid)) throw std::runtime_error("capture with id CaptureUid:" + c.first + " already exists");
    Capture& cap = addCapture(uid, c.second.second);
    if (c.second.first.size() < 6) throw std::runtime_error("inv_pose needs 6 values");
    for (size_t i = 0; i < 6; ++i) cap.inv_pose.params[i] = c.second.first[i];
  }
  for (auto& t : tags) {
    Aruco& a = addAruco(ArucoId(t.first));
    if (t.second.size() < 6) throw std::runtime_error("pose needs 6 values");
    for (size_t i = 0; i < 6; ++i) a.pose.params[i] = t.second[i];
  }
  for (auto& b : blks) {
    auto ci = capture_map_.find(CaptureUid(b.cap));
    auto ai = aruco_map_.find(ArucoId(b.tag));
    if (ci == capture_map_.end() || ai == aruco_map_.end()) throw std::out_of_range("block refers to an unknown capture or aruco");
    if (b.rect.size() != 8) throw std::runtime_error("aruco_rect has wrong number of values");
    ArucoRect r;
    for (unsigned i = 0; i < 4; ++i) { r.corners[i].x = b.rect[2 * i]; r.corners[i].y = b.rect[2 * i + 1]; }
    addBlock(r, ci->second, ai->second);
  }
  if (width < 0 || height < 0) throw std::runtime_error("camera section needs width and height");
  camera_.size = std::make_pair(width, height);
  for (size_t i = 0; i < cam_params.size() && i < 3; ++i) camera_.params[i] = cam_params[i];
}

void ArSlamSolver::saveYaml(std::ostream& out) const {
  auto seq = [](std::ostream& o, const double* v, size_t n) {
    o << "[";
    for (size_t i = 0; i < n; ++i) o << (i ? ", " : "") << num(v[i]);
    o << "]";
  };
  out << "blocks:";
  if (blocks_.empty()) out << "\n  []";
  for (const Block& b : blocks_) {
    double r[8];
    for (unsigned i = 0; i < 4; ++i) { r[2 * i] = b.aruco_rect.corners[i].x; r[2 * i + 1] = b.aruco_rect.corners[i].y; }
    out << "\n  - capture: " << at(b.capture).uid.uid << "\n    aruco: " << at(b.aruco).id.id << "\n    aruco_rect: ";
    seq(out, r, 8);
  }
  out << "\ncaptures:";
  if (captures_.empty()) out << "\n  {}";
  for (const Capture& c : captures_) {
    out << "\n  " << c.uid.uid << ":\n    inv_pose: ";
    seq(out, c.inv_pose.params.data(), 6);
    out << "\n    img_fn: " << c.img_fn;
  }
  out << "\narucos:";
  if (arucos_.empty()) out << "\n  {}";
  for (const Aruco& a : arucos_) {
    out << "\n  " << a.id.id << ":\n    pose: ";
    seq(out, a.pose.params.data(), 6);
  }
  out << "\ncamera:\n  params: ";
  seq(out, camera_.params.data(), 3);
  if (camera_.size.has_value()) out << "\n  width: " << camera_.size->first << "\n  height: " << camera_.size->second;
  out << std::endl;
  if (!out.good()) throw std::runtime_error("Yaml emit is not good");
}

void ArSlamSolver::printCameras() const {
  std::cout << "\tf=" << camera_.params[0] << "\tl1=" << camera_.params[1] << "\tl1=" << camera_.params[2] << std::endl;
}

void ArSlamSolver::compareProjections() const {
  // residual dump of every added block (reference :175-189, :576-589) from one GPU evaluation
  std::vector<BlockHandle> added;
  for (const Block& b : blocks_) if (b.added) added.push_back(b.handle);
  if (added.empty()) return;
  ArSlamSolver* self = const_cast<ArSlamSolver*>(this);
  std::vector<BlockHandle> keep = self->problem_blocks_;
  self->problem_blocks_ = added;
  std::vector<int32_t> ci, ti;
  std::vector<double> rect;
  for (BlockHandle h : added) {
    const Block& b = at(h);
    ci.push_back(b.capture.idx); ti.push_back(b.aruco.idx);
    for (const Point& p : b.aruco_rect.corners) { rect.push_back(p.x); rect.push_back(p.y); }
  }
  std::vector<double> cap(6 * captures_.size()), tag(6 * arucos_.size()), res(8 * added.size());
  for (size_t i = 0; i < captures_.size(); ++i) std::copy_n(captures_[i].data(), 6, cap.begin() + 6 * i);
  for (size_t i = 0; i < arucos_.size(); ++i) std::copy_n(arucos_[i].data(), 6, tag.begin() + 6 * i);
  arslam_solver* g = self->handle();
  check(g, arslam_set_problem(g, captures_.size(), arucos_.size(), added.size(), ci.data(), ti.data(), rect.data()), "set_problem");
  self->device_blocks_ = added;
  self->device_captures_ = captures_.size();
  self->device_arucos_ = arucos_.size();
  check(g, arslam_set_params(g, camera_.params.data(), cap.data(), tag.data()), "set_params");
  double cost = 0;
  check(g, arslam_evaluate(g, &cost, res.data(), nullptr, nullptr, nullptr), "evaluate");
  for (size_t k = 0; k < added.size(); ++k) {
    std::cout << "Block " << added[k].idx << std::endl;
    for (unsigned i = 0; i < 4; ++i) {
      const Point& c = at(added[k]).aruco_rect.corners[i];
      std::cout << "  point " << i << std::endl
                << "    x " << c.x << " dx " << res[8 * k + 2 * i] << std::endl
                << "    y " << c.y << " dy " << res[8 * k + 2 * i + 1] << std::endl;
    }
  }
  self->problem_blocks_ = keep;
}

std::optional<CaptureHandle> ArSlamSolver::addDetections(const ar_slam_interfaces::msg::Detections& d) {
  if (d.detections.empty()) return std::nullopt;
  const std::pair<int, int> image_size((int)d.image_width, (int)d.image_height);
  if (camera_.size.has_value()) {
    if (camera_.size.value() != image_size) {
      std::cerr << "WARN Mismatched image size expected [" << camera_.size->first << " x " << camera_.size->second
                << "] got [" << image_size.first << " x " << image_size.second << "]" << std::endl;
      return std::nullopt;
    }
  } else {
    camera_.size = image_size;
  }
  Capture& capture = addCapture(CaptureUid(d.capture_uid), d.image_path);  // throws on a duplicate uid (hpp:422-425)
  for (const auto& det : d.detections) {
    Aruco& aruco = getOrAddAruco(ArucoId(det.id));
    addBlock(ArucoRect{det}, capture.handle, aruco.handle);
  }
  unsolved_captures_.insert(capture.handle);
  return capture.handle;
}

static void rect_of(const Block& block, double rect[8]) {
  for (unsigned i = 0; i < 4; ++i) { rect[2 * i] = block.aruco_rect.corners[i].x; rect[2 * i + 1] = block.aruco_rect.corners[i].y; }
}

// initCapturePose (ar_slam_util.cpp:91-108): on the host arrays, or queued for the device while the
// parameters live there (the tag pose it reads is the previous solve's result, which only the GPU has)
void ArSlamSolver::seedCapture(Capture& capture, const Block& from_block) {
  double rect[8];
  rect_of(from_block, rect);
  if (host_stale_) {
    PendingSeed p{true, (int32_t)capture.handle.idx, (int32_t)from_block.aruco.idx, {}};
    std::copy_n(rect, 8, p.rect);
    pending_seeds_.push_back(p);
  } else {
    ars::seed_capture_pose(rect, camera_.params[0], at(from_block.aruco).data(), options_.tag_size, capture.data());
  }
}
// initArPose (ar_slam_util.cpp:111-128)
void ArSlamSolver::seedAruco(Aruco& aruco, const Capture& from_capture, const Block& block) {
  double rect[8];
  rect_of(block, rect);
  if (host_stale_) {
    PendingSeed p{false, (int32_t)aruco.handle.idx, (int32_t)from_capture.handle.idx, {}};
    std::copy_n(rect, 8, p.rect);
    pending_seeds_.push_back(p);
  } else {
    ars::seed_tag_pose(rect, camera_.params[0], from_capture.data(), options_.tag_size, aruco.data());
  }
}

void ArSlamSolver::syncFromDevice() {
  if (!host_stale_ || !gpu_) { host_stale_ = false; return; }
  std::vector<double> cap(6 * device_captures_), tag(6 * device_arucos_);
  check(gpu_, arslam_get_params(gpu_, camera_.params.data(), cap.data(), tag.data()), "get_params");
  for (size_t i = 0; i < device_captures_; ++i) std::copy_n(cap.begin() + 6 * i, 6, captures_[i].data());
  for (size_t i = 0; i < device_arucos_; ++i) std::copy_n(tag.begin() + 6 * i, 6, arucos_[i].data());
  host_stale_ = false;
}

void ArSlamSolver::addCaptureBlocksToProblem(Capture& capture) {
  for (BlockHandle bh : capture.blocks) {
    Block& block = at(bh);
    Aruco& aruco = at(block.aruco);
    if (!aruco.initialized) {
      aruco.initialized = true;
      seedAruco(aruco, capture, block);  // initArPose
    }
    if (block.added) throw std::runtime_error("block for capture was somehow already added?");
    block.added = true;
    problem_blocks_.push_back(bh);  // == problem_.AddResidualBlock(cost, nullptr, camera, capture, aruco)
  }
}

void ArSlamSolver::solveCapture(Capture& capture, std::optional<BlockHandle> init_block_handle) {
  if (init_block_handle.has_value()) seedCapture(capture, at(init_block_handle.value()));  // initCapturePose
  addCaptureBlocksToProblem(capture);
  optimize(capture);
}

void ArSlamSolver::solveIncremental() {
  if (unsolved_captures_.size() == captures_.size() && !unsolved_captures_.empty()) {
    CaptureHandle h = *unsolved_captures_.begin();
    std::cout << "Solving initial capture " << h.idx << std::endl;
    unsolved_captures_.erase(h);
    solveCapture(at(h), std::nullopt);
  }
  std::cout << "Solve incremental with " << unsolved_captures_.size() << " unsolved captures" << std::endl;
  bool repeat_solve = false;
  do {
    repeat_solve = false;
    for (auto itr = unsolved_captures_.begin(); itr != unsolved_captures_.end(); ++itr) {
      Capture& capture = at(*itr);
      for (BlockHandle bh : capture.blocks) {
        if (at(at(bh).aruco).initialized) {
          std::cout << "Capture CaptureUid:" << capture.uid.uid << " can be solved through ArucoId:" << at(at(bh).aruco).id.id << std::endl;
          repeat_solve = true;
          itr = unsolved_captures_.erase(itr);
          solveCapture(capture, bh);
          break;
        }
      }
      if (itr == unsolved_captures_.end()) break;
    }
  } while (repeat_solve);
  syncFromDevice();  // the callers (detection_callback: getTransforms / markers) read the host arrays
}

void ArSlamSolver::solve() {
  if (captures_.empty()) return;
  std::deque<CaptureHandle> open_captures;
  unsigned best_cap_idx = 0;
  {
    size_t best = captures_.front().blocks.size();
    for (unsigned i = 1; i < captures_.size(); ++i)
      if (captures_[i].blocks.size() > best) { best = captures_[i].blocks.size(); best_cap_idx = i; }
    std::cout << "using best capture " << best_cap_idx << " with " << best << " tags" << std::endl;
  }
  Capture& best_capture = captures_[best_cap_idx];
  best_capture.init_block = BlockHandle(~0u);  // prevents the capture from being queued again
  open_captures.emplace_back(best_capture.handle);
  while (!open_captures.empty()) {
    // captures_per_solve_ captures join the problem per optimize() call (1: the reference's schedule;
    // more: its TODO at ar_slam_util.cpp:810)
    Capture* last = nullptr;
    for (unsigned batch = 0; batch < std::max(1u, captures_per_solve_) && !open_captures.empty(); ++batch) {
      CaptureHandle h = open_captures.front();
      open_captures.pop_front();
      std::cout << "Processing capture " << h.idx << std::endl;
      Capture& capture = at(h);
      if (h.idx != best_cap_idx) seedCapture(capture, at(capture.init_block.value()));  // initCapturePose
      addCaptureBlocksToProblem(capture);
      addConnectedCaptures(capture, open_captures);  // structure only: the queue is the same as when it follows optimize()
      last = &capture;
    }
    optimize(*last);
  }
  syncFromDevice();
}

void ArSlamSolver::addConnectedCaptures(const Capture& base, std::deque<CaptureHandle>& open_captures) {
  for (BlockHandle bbh : base.blocks) {
    Aruco& base_aruco = at(at(bbh).aruco);
    for (BlockHandle bh : base_aruco.blocks) {
      Capture& capture = at(at(bh).capture);
      if (!capture.init_block.has_value()) {
        capture.init_block = bh;
        open_captures.emplace_back(capture.handle);
      }
    }
  }
}

// One batched GPU call for all new captures: they are mutually independent
// (tags and camera constant, ar_slam_util.cpp:965,972), so the reference's
// serial loop (:897-900) and the batch give the same poses.
void ArSlamSolver::localizeMany(unsigned first_loc_cap_idx) {
  const size_t n_loc = captures_.size() > first_loc_cap_idx ? captures_.size() - first_loc_cap_idx : 0;
  if (n_loc == 0) return;
  resetProblem();
  std::vector<int32_t> offsets(1, 0), tag_idx, seed(n_loc, -1);
  std::vector<double> rect, tag(6 * arucos_.size()), pose(6 * n_loc);
  for (size_t i = 0; i < arucos_.size(); ++i) std::copy_n(arucos_[i].data(), 6, tag.begin() + 6 * i);
  for (size_t i = 0; i < n_loc; ++i) {
    Capture& capture = captures_.at(first_loc_cap_idx + i);
    std::copy_n(capture.data(), 6, pose.begin() + 6 * i);
    int k = 0;
    for (BlockHandle cbh : capture.blocks) {
      Block& block = at(cbh);
      if (seed[i] < 0) {  // first tag shared with one of the "mapping" captures (:911-927)
        for (BlockHandle bh : at(block.aruco).blocks)
          if (at(bh).capture.idx < first_loc_cap_idx) { seed[i] = k; break; }
      }
      tag_idx.push_back(block.aruco.idx);
      for (const Point& p : block.aruco_rect.corners) { rect.push_back(p.x); rect.push_back(p.y); }
      ++k;
    }
    offsets.push_back((int32_t)tag_idx.size());
    if (seed[i] < 0) {
      std::cout << "WARNING : Cannot find connected ar tags for capture " << capture.handle.idx << std::endl;
    } else {
      for (BlockHandle cbh : capture.blocks) {
        if (at(cbh).added) throw std::runtime_error("block for capture was somehow already added?");
        at(cbh).added = true;
      }
    }
  }
  if (tag_idx.empty()) return;
  arslam_solver* g = handle();
  check(g, arslam_set_options(g, &options_), "set_options");
  check(g, arslam_localize_batch(g, n_loc, offsets.data(), tag_idx.data(), rect.data(), seed.data(), arucos_.size(),
                                 camera_.params.data(), tag.data(), pose.data(), nullptr, nullptr, nullptr),
        "localize_batch");
  for (size_t i = 0; i < n_loc; ++i)
    if (seed[i] >= 0) std::copy_n(pose.begin() + 6 * i, 6, captures_.at(first_loc_cap_idx + i).data());
}

// == ceres::Solve(options{max_num_iterations 50, DENSE_SCHUR}, &problem_, &summary)
void ArSlamSolver::optimize(const Capture&) {
  if (problem_blocks_.empty()) return;
  // The incremental schedules only ever append residual blocks between two resetProblem calls
  // (reference :723, :832 add to the live ceres::Problem).  When the blocks already on the GPU
  // are a prefix of the problem, only the new ones are uploaded (arslam_append_blocks); the
  // sorted views are rebuilt on the device either way.
  const bool is_extension = !device_blocks_.empty() && device_blocks_.size() < problem_blocks_.size() &&
                            device_captures_ <= captures_.size() && device_arucos_ <= arucos_.size() &&
                            std::equal(device_blocks_.begin(), device_blocks_.end(), problem_blocks_.begin(),
                                       [](BlockHandle a, BlockHandle b) { return a.idx == b.idx; });
  const bool is_same = device_blocks_.size() == problem_blocks_.size() && device_captures_ == captures_.size() &&
                       device_arucos_ == arucos_.size() &&
                       std::equal(device_blocks_.begin(), device_blocks_.end(), problem_blocks_.begin(),
                                  [](BlockHandle a, BlockHandle b) { return a.idx == b.idx; });
  const size_t first = is_extension ? device_blocks_.size() : 0;
  std::vector<int32_t> ci, ti;
  std::vector<double> rect;
  if (!is_same) {
    const size_t n_up = problem_blocks_.size() - first;
    ci.reserve(n_up); ti.reserve(n_up); rect.reserve(8 * n_up);
    for (size_t k = first; k < problem_blocks_.size(); ++k) {
      const Block& b = at(problem_blocks_[k]);
      ci.push_back(b.capture.idx);
      ti.push_back(b.aruco.idx);
      for (const Point& p : b.aruco_rect.corners) { rect.push_back(p.x); rect.push_back(p.y); }
    }
  }
  arslam_solver* g = handle();
  std::cout << "Starting solver..." << std::endl;
  check(g, arslam_set_options(g, &options_), "set_options");
  const bool on_device = device_resident_ && device_params_ && host_stale_ && is_extension;
  if (on_device) {
    // the parameters of everything solved so far are on the GPU: append the new blocks (parameters stay),
    // then seed the new captures / tags there, in the order the schedule asked for them
    check(g, arslam_append_blocks(g, captures_.size(), arucos_.size(), ci.size(), ci.data(), ti.data(), rect.data()), "append_blocks");
    size_t i = 0;
    while (i < pending_seeds_.size()) {
      size_t j = i;
      std::vector<int32_t> tgt, src;
      std::vector<double> rr;
      while (j < pending_seeds_.size() && pending_seeds_[j].is_capture == pending_seeds_[i].is_capture) {
        tgt.push_back(pending_seeds_[j].target); src.push_back(pending_seeds_[j].source);
        rr.insert(rr.end(), pending_seeds_[j].rect, pending_seeds_[j].rect + 8);
        ++j;
      }
      if (pending_seeds_[i].is_capture) check(g, arslam_seed_captures(g, tgt.size(), tgt.data(), src.data(), rr.data()), "seed_captures");
      else check(g, arslam_seed_tags(g, tgt.size(), tgt.data(), src.data(), rr.data()), "seed_tags");
      i = j;
    }
    pending_seeds_.clear();
  } else {
    if (host_stale_) syncFromDevice();
    if (!pending_seeds_.empty()) throw std::runtime_error("optimize: seeds were queued for the device but the problem is not an extension");
    std::vector<double> cap(6 * captures_.size()), tag(6 * arucos_.size());
    for (size_t i = 0; i < captures_.size(); ++i) std::copy_n(captures_[i].data(), 6, cap.begin() + 6 * i);
    for (size_t i = 0; i < arucos_.size(); ++i) std::copy_n(arucos_[i].data(), 6, tag.begin() + 6 * i);
    if (is_extension)
      check(g, arslam_append_blocks(g, captures_.size(), arucos_.size(), ci.size(), ci.data(), ti.data(), rect.data()), "append_blocks");
    else if (!is_same)
      check(g, arslam_set_problem(g, captures_.size(), arucos_.size(), ci.size(), ci.data(), ti.data(), rect.data()), "set_problem");
    check(g, arslam_set_params(g, camera_.params.data(), cap.data(), tag.data()), "set_params");
  }
  device_blocks_ = problem_blocks_;
  device_captures_ = captures_.size();
  device_arucos_ = arucos_.size();
  // like the reference, a solver that does not converge is not an error (summary discarded at :1013-1017);
  // API misuse and CUDA failures are
  check(g, arslam_solve(g, &last_summary_, nullptr, 0), "solve");
  summaries_.push_back(last_summary_);
  if (device_resident_) {
    // only the intrinsics come back (the seeds of the next capture need the focal length; it is mirrored on
    // the host by the library, so this moves nothing); the poses stay on the GPU until the schedule returns
    check(g, arslam_get_params(g, camera_.params.data(), nullptr, nullptr), "get_params");
    device_params_ = true;
    host_stale_ = true;
  } else {
    std::vector<double> cap(6 * captures_.size()), tag(6 * arucos_.size());
    check(g, arslam_get_params(g, camera_.params.data(), cap.data(), tag.data()), "get_params");
    for (size_t i = 0; i < captures_.size(); ++i) std::copy_n(cap.begin() + 6 * i, 6, captures_[i].data());
    for (size_t i = 0; i < arucos_.size(); ++i) std::copy_n(tag.begin() + 6 * i, 6, arucos_[i].data());
  }
}

void ArSlamSolver::resetProblem() { problem_blocks_.clear(); }

// ---------------------------------------------------------------- ROS outputs ---
// What ArSlam::detection_callback publishes after every solveIncremental (reference
// ar_slam/src/ar_slam.cpp:133 tf, :145 CameraInfo, :154 markers).  Same conventions as
// ar_slam_util.cpp:1027-1162: tags are forward poses (world <- tag: translation and rotation as
// stored), captures are INVERSE poses (p_cam = R(w)(p_world + t)): the published world <- camera
// transform negates both the translation and the angle-axis vector; quaternions leave Ceres'
// AngleAxisToQuaternion as (w, x, y, z) and are written into the message's x, y, z, w fields.
namespace {
void angle_axis_to_quaternion(const double aa[3], double q[4]) {  // ceres::AngleAxisToQuaternion, w first
  ars::aa_to_quat(aa, q);
}
template <typename Stamp, typename T> void set_stamp(Stamp& dst, const T& src) { dst = src; }
}  // namespace

std::vector<geometry_msgs::msg::TransformStamped> ArSlamSolver::getTransforms(const arslam_ros::Time& stamp) const {
  std::vector<geometry_msgs::msg::TransformStamped> transforms;
  transforms.reserve(captures_.size() + arucos_.size());
  for (const Aruco& aruco : arucos_) {
    transforms.emplace_back();
    auto& t = transforms.back();
    set_stamp(t.header.stamp, stamp);
    t.header.frame_id = "world";
    t.child_frame_id = aruco.id.id;
    const auto& pose = aruco.pose.params;
    t.transform.translation.x = pose[0];
    t.transform.translation.y = pose[1];
    t.transform.translation.z = pose[2];
    double q[4];
    angle_axis_to_quaternion(&pose[3], q);
    t.transform.rotation.w = q[0];
    t.transform.rotation.x = q[1];
    t.transform.rotation.y = q[2];
    t.transform.rotation.z = q[3];
  }
  for (const Capture& capture : captures_) {
    transforms.emplace_back();
    auto& t = transforms.back();
    set_stamp(t.header.stamp, stamp);
    t.header.frame_id = "world";
    t.child_frame_id = capture.uid.uid;
    const auto& inv_pose = capture.inv_pose.params;
    const double rot[3] = {-inv_pose[3], -inv_pose[4], -inv_pose[5]};
    double q[4];
    angle_axis_to_quaternion(rot, q);
    t.transform.rotation.w = q[0];
    t.transform.rotation.x = q[1];
    t.transform.rotation.y = q[2];
    t.transform.rotation.z = q[3];
    t.transform.translation.x = -inv_pose[0];
    t.transform.translation.y = -inv_pose[1];
    t.transform.translation.z = -inv_pose[2];
  }
  return transforms;
}

sensor_msgs::msg::CameraInfo ArSlamSolver::getCameraInfo() const {
  sensor_msgs::msg::CameraInfo info;
  info.distortion_model = sensor_msgs::distortion_models::PLUMB_BOB;
  info.d = {0, 0, 0, 0, 0};  // plumb_bob: k1, k2, t1, t2, k3 -- the live model has no distortion
  const double fx = camera_.params[0], fy = camera_.params[0];
  if (!camera_.size.has_value()) throw std::runtime_error("getCameraInfo: image size unknown (no detections yet)");
  const double cx = camera_.size->first * 0.5, cy = camera_.size->second * 0.5;  // width, height
  info.k = {fx, 0.0, cx, 0.0, fy, cy, 0.0, 0.0, 1.0};
  info.r = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
  info.p = {fx, 0.0, cx, 0.0, 0.0, fy, cy, 0.0, 0.0, 0.0, 1.0, 0.0};
  return info;
}

void ArSlamSolver::appendArucoMarkers(std::vector<visualization_msgs::msg::Marker>& markers, arslam_ros::Time stamp) const {
  markers.reserve(markers.size() + arucos_.size() + 1);
  {
    markers.emplace_back();
    auto& marker = markers.back();
    set_stamp(marker.header.stamp, stamp);
    marker.action = marker.DELETEALL;
    marker.ns = "arucos";
  }
  for (unsigned idx = 0; idx < arucos_.size(); ++idx) {
    const Aruco& aruco = arucos_[idx];
    markers.emplace_back();
    auto& marker = markers.back();
    set_stamp(marker.header.stamp, stamp);
    marker.header.frame_id = aruco.id.id;
    marker.type = marker.CUBE;
    marker.action = marker.ADD;
    marker.ns = "arucos";
    marker.id = idx;
    marker.scale.x = aruco_size;
    marker.scale.y = aruco_size;
    marker.scale.z = 0.01;  // 1 cm thick
    marker.color.a = 0.8;
    marker.color.r = 1.0;
    marker.frame_locked = true;
  }
}
