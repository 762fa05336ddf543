// bin/chan_vese -- C++14 host front-end with the reference's command-line surface over the C ABI
// (include/chan_vese_b200.h).  Mirrors the option table of /root/reference/src/main.cpp:756-781 (same long and
// short names, defaults and validation messages, :786-869), the "_pm" / "_selection" output naming (:158-167,
// :946, :1005) and the silent stdout.  All numerics run in libchan_vese_b200.so; this file only parses options,
// reads/writes images and composites the selection.
//
// Host image I/O: OpenCV's C++ libraries (which the reference uses for imread/imwrite) and Boost are not
// installed in this image, so files are binary PNM (P6 colour, P5 gray); the two functions read_image/write_image
// are the only place an OpenCV-linked build would change.  The interactive -R/-C contour selection (a GUI window,
// src/main.cpp:899-921) becomes --rect x,y,w,h / --circ cx,cy,r with the same level sets
// (InteractiveDataRect.cpp:20-27, InteractiveDataCirc.cpp:18-25).  -V: the reference writes "<stem>.avi" (XVID through
// highgui, src/VideoWriterManager.cpp:24-54) with frame 0 = the initial contour and one frame per step (src/main.cpp:929,
// :997); without a video encoder on this host the same frames go, uncompressed, into "<stem>.ppms" -- binary PPM images
// back to back (`ffmpeg -f image2pipe -vcodec ppm -r <fps> -i <stem>.ppms out.avi` encodes them) -- fed by the
// asynchronous per-step mask observer of the ABI (cvb_csv_run_masks, CVB_MASK_CONTOUR = VideoWriterManager's threshold
// saturate_cast<uchar>(u) > 0, :65-68); the contour of the FINAL level set is also written as "<stem>_contour<ext>".
// Overlay text (-O) needs a font renderer and is not drawn.
#include <sys/stat.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "chan_vese_b200.h"

namespace {

[[noreturn]] void msg_exit(const std::string &msg) {  // src/main.cpp:173-178
    std::fprintf(stderr, "\n%s\n\n", msg.c_str());
    std::exit(EXIT_FAILURE);
}

std::string add_suffix(const std::string &path, const std::string &suffix) {  // src/main.cpp:158-167
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
    const std::string stem = has_ext ? path.substr(0, dot) : path;
    const std::string ext = has_ext ? path.substr(dot) : "";
    return stem + "_" + suffix + ext;
}

struct Image {
    int h = 0, w = 0, n = 0;                  // n = 1 (gray) or 3
    std::vector<std::vector<uint8_t>> planes;  // B,G,R order for colour (what cv::imread + cv::split give)
};

int pnm_int(FILE *f) {
    int c = std::fgetc(f);
    for (;;) {
        while (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = std::fgetc(f);
        if (c != '#') break;
        while (c != '\n' && c != EOF) c = std::fgetc(f);
    }
    if (c < '0' || c > '9') throw std::runtime_error("malformed PNM header");
    int v = 0;
    while (c >= '0' && c <= '9') {
        v = v * 10 + (c - '0');
        c = std::fgetc(f);
    }
    return v;
}

// read_image: the stand-in for cv::imread(path, grayscale ? GRAYSCALE : COLOR), src/main.cpp:877-881
Image read_image(const std::string &path, bool grayscale) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) msg_exit("Error on opening \"" + path + "\" (probably not an image)!");
    Image img;
    try {
        if (std::fgetc(f) != 'P') throw std::runtime_error("not a PNM file");
        const int kind = std::fgetc(f);
        if (kind != '5' && kind != '6') throw std::runtime_error("only binary PGM (P5) / PPM (P6) are supported");
        img.w = pnm_int(f);
        img.h = pnm_int(f);
        if (pnm_int(f) != 255 || img.w <= 0 || img.h <= 0) throw std::runtime_error("only 8-bit PNM is supported");
        const int src_n = kind == '6' ? 3 : 1;
        std::vector<uint8_t> raw((size_t)img.w * img.h * src_n);
        if (std::fread(raw.data(), 1, raw.size(), f) != raw.size()) throw std::runtime_error("truncated PNM data");
        const size_t np = (size_t)img.w * img.h;
        if (grayscale) {  // cv::imread's BGR -> gray weights (0.114 B + 0.587 G + 0.299 R), fixed-point as OpenCV does
            img.n = 1;
            img.planes.assign(1, std::vector<uint8_t>(np));
            for (size_t p = 0; p < np; ++p) {
                if (src_n == 1) {
                    img.planes[0][p] = raw[p];
                } else {
                    const int r = raw[3 * p], g = raw[3 * p + 1], b = raw[3 * p + 2];
                    img.planes[0][p] = (uint8_t)((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14);
                }
            }
        } else {
            img.n = 3;
            img.planes.assign(3, std::vector<uint8_t>(np));
            for (size_t p = 0; p < np; ++p)
                for (int k = 0; k < 3; ++k) img.planes[k][p] = src_n == 3 ? raw[3 * p + (2 - k)] : raw[p];
        }
    } catch (const std::exception &e) {
        std::fclose(f);
        msg_exit("Error on opening \"" + path + "\" (" + e.what() + ")!");
    }
    std::fclose(f);
    return img;
}

// write_image: the stand-in for cv::imwrite; planes in B,G,R order (or one gray plane)
void write_image(const std::string &path, const std::vector<const uint8_t *> &planes, int h, int w) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) msg_exit("Error: cannot write \"" + path + "\"!");
    const size_t np = (size_t)h * w;
    if (planes.size() == 1) {
        std::fprintf(f, "P5\n%d %d\n255\n", w, h);
        std::fwrite(planes[0], 1, np, f);
    } else {
        std::fprintf(f, "P6\n%d %d\n255\n", w, h);
        std::vector<uint8_t> row((size_t)w * 3);
        for (int i = 0; i < h; ++i) {
            for (int j = 0; j < w; ++j)
                for (int k = 0; k < 3; ++k) row[3 * j + k] = planes[2 - k][(size_t)i * w + j];
            std::fwrite(row.data(), 1, row.size(), f);
        }
    }
    std::fclose(f);
}

struct Options {
    std::string input;
    double mu = 0.5, nu = 0, dt = 1, eps = 1, tol = 0.001, fps = 10, K = 10, L = 0.25, T = 20;
    std::vector<double> lambda1, lambda2;
    int max_steps = -1;
    std::string text_position = "TL", line_color = "blue";
    bool segment = false, grayscale = false, video = false, overlay = false, invert = false, select = false;
    bool rectangle = false, circle = false, help = false, stats = false;
    int rect[4] = {0, 0, 0, 0}, circ[3] = {0, 0, 0};
    bool seen_dt = false, seen_mu = false, seen_l1 = false, seen_l2 = false, seen_L = false, seen_T = false;
};

const char *kHelp =
    "Allowed options:\n"
    "  -h [ --help ]                   this message\n"
    "  -i [ --input ] arg              input image (binary PGM/PPM)\n"
    "  --mu arg (=0.5)                 length penalty parameter (must be positive or zero)\n"
    "  --nu arg (=0)                   area penalty parameter\n"
    "  --dt arg (=1)                   timestep\n"
    "  --lambda1 arg                   penalty of variance inside the contour (default: 1's)\n"
    "  --lambda2 arg                   penalty of variance outside the contour (default: 1's)\n"
    "  -e [ --epsilon ] arg (=1)       smoothing parameter in Heaviside/delta\n"
    "  -t [ --tolerance ] arg (=0.001) tolerance in stopping condition\n"
    "  -N [ --max-steps ] arg (=-1)    maximum nof iterations (negative means unlimited)\n"
    "  -f [ --fps ] arg (=10)          video fps (accepted; the frame stream <stem>.ppms carries no timing)\n"
    "  -P [ --overlay-pos ] arg (=TL)  overlay tex position; allowed only: TL, BL, TR, BR\n"
    "  -l [ --line-color ] arg (=blue) contour color (allowed only: black, white, R, G, B, Y, M, C\n"
    "  -K [ --edge-coef ] arg (=10)    coefficient for enhancing edge detection in Perona-Malik\n"
    "  -L [ --laplacian-coef ] arg (=0.25) coefficient in the gradient FD scheme of Perona-Malik (must be [0, 1/4])\n"
    "  -T [ --segment-time ] arg (=20) number of smoothing steps in Perona-Malik\n"
    "  -S [ --segment ]                segment the image with Perona-Malik beforehand\n"
    "  -g [ --grayscale ]              read in as grayscale\n"
    "  -V [ --video ]                  per-step contour frames as <stem>.ppms (PPM stream) + the final contour ('_contour')\n"
    "  -O [ --overlay-text ]           add overlay text (accepted, unused)\n"
    "  -I [ --invert-selection ]       invert selected region (see: select)\n"
    "  -s [ --select ]                 separate the region encolosed by the contour (adds suffix '_selection')\n"
    "  -R [ --rectangle ]              rectangular contour; give it with --rect x,y,w,h (no GUI on this host)\n"
    "  -C [ --circle ]                 circular contour; give it with --circ cx,cy,r (no GUI on this host)\n"
    "  --rect x,y,w,h                  rectangular initial contour (implies -R)\n"
    "  --circ cx,cy,r                  circular initial contour (implies -C)\n"
    "  --stats                         print steps, norm and timings to stderr (stdout stays silent)\n";

bool is_number(const char *s) {
    char *end = nullptr;
    std::strtod(s, &end);
    return end != s && *end == '\0';
}
double to_double(const std::string &opt, const char *s) {
    if (!is_number(s)) throw std::runtime_error("the argument ('" + std::string(s) + "') for option '--" + opt + "' is invalid");
    return std::strtod(s, nullptr);
}
void parse_list(const std::string &opt, const char *s, int *out, int n) {
    std::string v(s);
    size_t pos = 0;
    for (int k = 0; k < n; ++k) {
        const size_t comma = v.find(',', pos);
        const std::string tok = v.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        if (tok.empty() || (k < n - 1 && comma == std::string::npos))
            throw std::runtime_error("option '--" + opt + "' needs " + std::to_string(n) + " comma-separated integers");
        out[k] = std::atoi(tok.c_str());
        pos = comma + 1;
    }
}

// Boost.ProgramOptions-compatible subset: long/short names, "--opt value" and "--opt=value", bool switches,
// multitoken --lambda1/--lambda2, negative numbers accepted as values (README.md:60 uses --nu -293).
Options parse(int argc, char **argv) {
    Options o;
    struct Spec { const char *lname; char sname; int kind; };  // kind 0 switch, 1 value, 2 multitoken
    static const Spec specs[] = {
        {"help", 'h', 0}, {"input", 'i', 1}, {"mu", 0, 1}, {"nu", 0, 1}, {"dt", 0, 1}, {"lambda1", 0, 2}, {"lambda2", 0, 2},
        {"epsilon", 'e', 1}, {"tolerance", 't', 1}, {"max-steps", 'N', 1}, {"fps", 'f', 1}, {"overlay-pos", 'P', 1},
        {"line-color", 'l', 1}, {"edge-coef", 'K', 1}, {"laplacian-coef", 'L', 1}, {"segment-time", 'T', 1},
        {"segment", 'S', 0}, {"grayscale", 'g', 0}, {"video", 'V', 0}, {"overlay-text", 'O', 0},
        {"invert-selection", 'I', 0}, {"select", 's', 0}, {"rectangle", 'R', 0}, {"circle", 'C', 0}, {"rect", 0, 1},
        {"circ", 0, 1}, {"stats", 0, 0}};
    std::vector<std::string> args(argv + 1, argv + argc);
    for (size_t a = 0; a < args.size(); ++a) {
        std::string tok = args[a], inline_val;
        bool has_inline = false;
        const Spec *sp = nullptr;
        std::vector<std::pair<const Spec *, bool>> shorts;  // bundled short switches, e.g. -sg
        if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
            std::string name = tok.substr(2);
            const size_t eq = name.find('=');
            if (eq != std::string::npos) {
                inline_val = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_inline = true;
            }
            for (const auto &s : specs)
                if (name == s.lname) sp = &s;
            if (!sp) throw std::runtime_error("unrecognised option '--" + name + "'");
        } else if (tok.size() >= 2 && tok[0] == '-' && !is_number(tok.c_str())) {
            for (size_t c = 1; c < tok.size(); ++c) {
                const Spec *f = nullptr;
                for (const auto &s : specs)
                    if (s.sname && s.sname == tok[c]) f = &s;
                if (!f) throw std::runtime_error(std::string("unrecognised option '-") + tok[c] + "'");
                if (f->kind != 0) {  // a value option ends the bundle; the rest of the token is its value
                    sp = f;
                    if (c + 1 < tok.size()) {
                        inline_val = tok.substr(c + 1);
                        has_inline = true;
                    }
                    break;
                }
                shorts.push_back({f, true});
            }
        } else {
            throw std::runtime_error("too many positional options have been specified on the command line");
        }
        auto set_switch = [&](const Spec *s) {
            const std::string n = s->lname;
            if (n == "help") o.help = true;
            else if (n == "segment") o.segment = true;
            else if (n == "grayscale") o.grayscale = true;
            else if (n == "video") o.video = true;
            else if (n == "overlay-text") o.overlay = true;
            else if (n == "invert-selection") o.invert = true;
            else if (n == "select") o.select = true;
            else if (n == "rectangle") o.rectangle = true;
            else if (n == "circle") o.circle = true;
            else if (n == "stats") o.stats = true;
        };
        for (auto &s : shorts) set_switch(s.first);
        if (!sp) continue;
        if (sp->kind == 0) {
            set_switch(sp);
            continue;
        }
        const std::string n = sp->lname;
        std::vector<std::string> vals;
        if (has_inline) vals.push_back(inline_val);
        if (sp->kind == 2) {  // multitoken: take following tokens while they look like values
            while (a + 1 < args.size() && is_number(args[a + 1].c_str())) vals.push_back(args[++a]);
        } else if (!has_inline) {
            if (a + 1 >= args.size()) throw std::runtime_error("the required argument for option '--" + n + "' is missing");
            vals.push_back(args[++a]);
        }
        if (vals.empty()) throw std::runtime_error("the required argument for option '--" + n + "' is missing");
        const char *v = vals[0].c_str();
        if (n == "input") o.input = v;
        else if (n == "mu") { o.mu = to_double(n, v); o.seen_mu = true; }
        else if (n == "nu") o.nu = to_double(n, v);
        else if (n == "dt") { o.dt = to_double(n, v); o.seen_dt = true; }
        else if (n == "epsilon") o.eps = to_double(n, v);
        else if (n == "tolerance") o.tol = to_double(n, v);
        else if (n == "max-steps") o.max_steps = (int)to_double(n, v);
        else if (n == "fps") o.fps = to_double(n, v);
        else if (n == "overlay-pos") o.text_position = v;
        else if (n == "line-color") o.line_color = v;
        else if (n == "edge-coef") o.K = to_double(n, v);
        else if (n == "laplacian-coef") { o.L = to_double(n, v); o.seen_L = true; }
        else if (n == "segment-time") { o.T = to_double(n, v); o.seen_T = true; }
        else if (n == "rect") { parse_list(n, v, o.rect, 4); o.rectangle = true; }
        else if (n == "circ") { parse_list(n, v, o.circ, 3); o.circle = true; }
        else if (n == "lambda1" || n == "lambda2") {
            std::vector<double> &dst = n == "lambda1" ? o.lambda1 : o.lambda2;
            for (const auto &s : vals) dst.push_back(to_double(n, s.c_str()));
            (n == "lambda1" ? o.seen_l1 : o.seen_l2) = true;
        }
    }
    return o;
}

bool iequals(const std::string &a, const char *b) {
    if (a.size() != std::strlen(b)) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}

void validate(Options &o, uint8_t color[3]) {  // src/main.cpp:786-869, same messages
    struct stat st;
    if (o.input.empty()) msg_exit("Error: you have to specify input file name!");
    if (stat(o.input.c_str(), &st) != 0) msg_exit("Error: file \"" + o.input + "\" does not exists!");
    if (o.seen_dt && o.dt <= 0) msg_exit("Cannot have negative or zero timestep: " + std::to_string(o.dt) + ".");
    if (o.seen_mu && o.mu < 0) msg_exit("Length penalty parameter cannot be negative: " + std::to_string(o.mu) + ".");
    auto check_lambda = [&](std::vector<double> &l, bool seen, const char *name) {
        const std::string nm(name);
        if (seen) {
            if (o.grayscale && l.size() != 1) msg_exit("Too many " + nm + " values for a grayscale image.");
            if (!o.grayscale && l.size() != 3) msg_exit("Number of " + nm + " values must be 3 for a colored input image.");
            for (double v : l)
                if (v < 0) msg_exit(o.grayscale ? "The value of " + nm + " cannot be negative." : "Any value of " + nm + " cannot be negative.");
        } else {
            l.assign(o.grayscale ? 1 : 3, 1.0);
        }
    };
    check_lambda(o.lambda1, o.seen_l1, "lambda1");
    check_lambda(o.lambda2, o.seen_l2, "lambda2");
    if (!(iequals(o.text_position, "TL") || iequals(o.text_position, "BL") || iequals(o.text_position, "TR") ||
          iequals(o.text_position, "BR")))
        msg_exit("Invalid text position requested.\nCorrect values are: TL -- top left\n                    BL -- bottom left\n"
                 "                    TR -- top right\n                    BR -- bottom right");
    struct { const char *name; uint8_t bgr[3]; } colors[] = {  // ChanVese::Colors, src/main.cpp:111-118 (B,G,R)
        {"red", {0, 0, 255}}, {"green", {0, 255, 0}}, {"blue", {255, 0, 0}}, {"black", {0, 0, 0}},
        {"white", {255, 255, 255}}, {"magenta", {255, 0, 255}}, {"yellow", {0, 255, 255}}, {"cyan", {255, 255, 0}}};
    bool found = false;
    for (auto &c : colors)
        if (iequals(o.line_color, c.name)) {
            std::memcpy(color, c.bgr, 3);
            found = true;
        }
    if (!found)
        msg_exit("Invalid contour color requested.\nCorrect values are: red, green, blue, black, white, magenta, yellow, cyan.");
    if (o.seen_L && (o.L > 0.25 || o.L < 0))
        msg_exit("The Laplacian coefficient in Perona-Malik segmentation must be between 0 and 0.25.");
    if (o.seen_T && o.T < o.L)
        msg_exit("The segmentation duration must exceed the value of Laplacian coefficient, " + std::to_string(o.L) + ".");
    if (o.rectangle && o.circle) msg_exit("Cannot initialize with both rectangular and circular contour");
    if (o.rectangle && o.rect[2] <= 0) msg_exit("No GUI on this host: give the rectangle as --rect x,y,w,h.");
    if (o.circle && o.circ[2] <= 0) msg_exit("No GUI on this host: give the circle as --circ cx,cy,r.");
    if (o.eps <= 0) msg_exit("Cannot have negative or zero smoothing parameter: " + std::to_string(o.eps) + ".");
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    uint8_t color[3] = {255, 0, 0};
    try {
        o = parse(argc, argv);
    } catch (const std::exception &e) {
        msg_exit("error: " + std::string(e.what()));  // src/main.cpp:871-874
    }
    if (o.help) {
        std::printf("%s\n", kHelp);
        return EXIT_SUCCESS;
    }
    validate(o, color);

    const Image img = read_image(o.input, o.grayscale);
    const int h = img.h, w = img.w, n = img.n;
    const size_t np = (size_t)h * w;
    if (o.max_steps < 0) o.max_steps = -1;  // unlimited, src/main.cpp:890

    // initial level set (src/main.cpp:897-923)
    std::vector<double> u(np);
    if (o.rectangle)
        cvb_levelset_rect(h, w, o.rect[0], o.rect[1], o.rect[2], o.rect[3], u.data());
    else if (o.circle)
        cvb_levelset_circ(h, w, o.circ[0], o.circ[1], o.circ[2], u.data());
    else
        cvb_levelset_checkerboard(h, w, u.data());

    cvb_context *ctx = nullptr;
    if (cvb_context_create(0, nullptr, &ctx) != CVB_OK) msg_exit(std::string("Error: ") + cvb_last_error(nullptr));
    auto check = [&](cvb_status st) {
        if (st != CVB_OK) msg_exit(std::string("Error: ") + cvb_last_error(ctx));
    };

    cvb_csv_params p{};
    p.mu = o.mu;
    p.nu = o.nu;
    p.dt = o.dt;
    p.eps = o.eps;
    for (int k = 0; k < 3; ++k) {
        p.lambda1[k] = k < n ? o.lambda1[k] : 1.0;
        p.lambda2[k] = k < n ? o.lambda2[k] : 1.0;
    }
    std::vector<const uint8_t *> planes;
    for (const auto &pl : img.planes) planes.push_back(pl.data());
    std::vector<std::vector<uint8_t>> pm(n, std::vector<uint8_t>(np));
    std::vector<uint8_t *> pm_ptrs;
    for (auto &pl : pm) pm_ptrs.push_back(pl.data());
    std::vector<uint8_t> mask(np);
    int steps = 0;
    double norm = 0;
    // contour of a 0/1 mask over the ORIGINAL image (VideoWriterManager draws over img, :43-45): mask pixels with a
    // 4-neighbour outside the mask take the contour colour
    auto draw = [&](const std::vector<uint8_t> &m01, std::vector<std::vector<uint8_t>> &fr) {
        for (int k = 0; k < 3; ++k) fr[k] = img.planes[n == 3 ? k : 0];
        auto in = [&](int i, int j) { return i >= 0 && i < h && j >= 0 && j < w && m01[(size_t)i * w + j]; };
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j)
                if (in(i, j) && !(in(i - 1, j) && in(i + 1, j) && in(i, j - 1) && in(i, j + 1)))
                    for (int k = 0; k < 3; ++k) fr[k][(size_t)i * w + j] = color[k];
    };
    if (!o.video) {
        // PM (optional) + the time-step loop + separate()'s mask in one resident pass (src/main.cpp:939-1001)
        check(cvb_segment(ctx, planes.data(), n, h, w, u.data(), o.segment ? 1 : 0, o.K, o.L, o.T, pm_ptrs.data(), &p, o.tol,
                          o.max_steps, &steps, &norm, o.invert ? 1 : 0, mask.data()));
    } else {
        // the same with the per-step frame stream: frame 0 = the initial contour (:929), then one frame per step (:997)
        struct Stream {
            FILE *f;
            int h, w, frames;
            std::vector<uint8_t> m01;
            std::vector<std::vector<uint8_t>> fr;
            decltype(draw) *draw_fn;
        } vs{nullptr, h, w, 0, std::vector<uint8_t>(np), std::vector<std::vector<uint8_t>>(3), &draw};
        const size_t dot = o.input.find_last_of('.'), slash = o.input.find_last_of('/');
        const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
        const std::string vname = (has_ext ? o.input.substr(0, dot) : o.input) + ".ppms";
        vs.f = std::fopen(vname.c_str(), "wb");
        if (!vs.f) msg_exit("Error: cannot open \"" + vname + "\" for writing");
        auto put = [](Stream &v) {
            (*v.draw_fn)(v.m01, v.fr);
            std::fprintf(v.f, "P6\n%d %d\n255\n", v.w, v.h);
            std::vector<uint8_t> rgb((size_t)v.h * v.w * 3);
            for (size_t q = 0; q < (size_t)v.h * v.w; ++q) {
                rgb[3 * q] = v.fr[2][q];
                rgb[3 * q + 1] = v.fr[1][q];
                rgb[3 * q + 2] = v.fr[0][q];
            }
            std::fwrite(rgb.data(), 1, rgb.size(), v.f);
            ++v.frames;
        };
        for (size_t q = 0; q < np; ++q) vs.m01[q] = std::nearbyint(u[q]) > 0;  // saturate_cast<uchar>(u) > 0, :65
        put(vs);
        static auto put_fn = +put;
        auto on_mask = [](const uint8_t *bits, int hh, int ww, int /*step*/, void *user) -> int {
            Stream &v = *static_cast<Stream *>(user);
            const int wb = (ww + 7) / 8;
            for (int i = 0; i < hh; ++i)
                for (int j = 0; j < ww; ++j) v.m01[(size_t)i * ww + j] = (bits[(size_t)i * wb + j / 8] >> (7 - j % 8)) & 1;
            put_fn(v);
            return 0;
        };
        const uint8_t *const *cur = planes.data();
        std::vector<const uint8_t *> pm_c;
        if (o.segment) {
            check(cvb_perona_malik(ctx, planes.data(), n, h, w, o.K, o.L, o.T, pm_ptrs.data(), nullptr));
            for (auto &pl : pm) pm_c.push_back(pl.data());
            cur = pm_c.data();
        }
        check(cvb_csv_run_masks(ctx, cur, n, h, w, u.data(), &p, o.tol, o.max_steps, &steps, &norm, CVB_MASK_CONTOUR, +on_mask, &vs));
        std::fclose(vs.f);
        check(cvb_mask(ctx, u.data(), h, w, o.invert ? 1 : 0, mask.data()));
    }
    if (o.segment) {  // cv::imwrite(add_suffix(input_filename, "pm"), smoothed_img), :946
        std::vector<const uint8_t *> out;
        for (auto &pl : pm) out.push_back(pl.data());
        write_image(add_suffix(o.input, "pm"), out, h, w);
    }
    if (o.select) {  // separate(): white canvas, original pixels under the mask, :386-405 (always 3 channels)
        std::vector<std::vector<uint8_t>> sel(3, std::vector<uint8_t>(np, 255));
        for (size_t q = 0; q < np; ++q)
            if (mask[q])
                for (int k = 0; k < 3; ++k) sel[k][q] = img.planes[n == 3 ? k : 0][q];
        write_image(add_suffix(o.input, "selection"), {sel[0].data(), sel[1].data(), sel[2].data()}, h, w);
    }
    if (o.video) {  // contour of the final level set: pixels of {u > 0.5} with a 4-neighbour outside it
        std::vector<std::vector<uint8_t>> fr(3);
        std::vector<uint8_t> m01(np);
        for (size_t q = 0; q < np; ++q) m01[q] = std::nearbyint(u[q]) > 0;
        draw(m01, fr);
        write_image(add_suffix(o.input, "contour"), {fr[0].data(), fr[1].data(), fr[2].data()}, h, w);
    }
    if (o.stats) {
        cvb_stats st{};
        cvb_context_get_stats(ctx, &st);
        std::fprintf(stderr, "steps=%d norm=%.17g pm_steps=%llu pm_ms=%.3f csv_ms=%.3f kernel_launches=%llu\n", steps, norm,
                     (unsigned long long)st.pm_step_launches, st.pm_ms, st.csv_ms, (unsigned long long)st.kernel_launches);
    }
    cvb_context_destroy(ctx);
    return EXIT_SUCCESS;
}
