// cli.cpp -- nnue_data_compression-compatible command line on top of libnnuepack.so.
//
// Mirrors readArgs / run / convert / compress / decompress / help of the reference
// (src/compress_file.cpp:1535-1709): same dispatch by file extension, same flags (including the
// quirk that "--append" is stored as "-append" and therefore ignored), same messages, same exit
// codes (0 everywhere except the argument-count error). The conversions themselves run on the
// GPU through the C ABI; this file only moves files in and out of host memory.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nnuepack.h"

namespace {

const std::string plainExtension = ".plain";
const std::string binExtension = ".bin";
const std::string binpackExtension = ".binpack";

typedef int (*driver_fn)(const void*, size_t, void*, size_t, size_t*);

bool endsWith(const std::string& str, const std::string& suffix)
{
    return str.size() >= suffix.size() && 0 == str.compare(str.size() - suffix.size(), suffix.size(), suffix);
}

bool fileExists(const std::string& name)
{
    std::ifstream f(name);
    return f.good();
}

bool isReferenceError(int rc) { return rc == NNP_ERR_BAD_MAGIC || rc == NNP_ERR_CHUNK_TOO_LARGE || rc == NNP_ERR_BAD_SFEN; }

// anything that is not one of the reference's own errors ends the process with a non-zero exit code
[[noreturn]] void fail(int rc)
{
    std::cerr << nnp_strerror(rc) << " " << nnp_last_cuda_error() << "\n";
    nnp_shutdown();
    std::exit(2);
}

struct HostBuffer {
    char* p = nullptr;
    explicit HostBuffer(size_t n) : p(static_cast<char*>(nnp_host_alloc(n + 1))) {}
    ~HostBuffer() { nnp_host_free(p); }
    HostBuffer(const HostBuffer&) = delete;
    HostBuffer& operator=(const HostBuffer&) = delete;
};

// The reference's progress lines (compress_file.cpp:1280-1283, :1318-1325, :1369-1372, :1395-1410): the
// compressors report every 100 000 positions, the decompressors at every flush of their 1 MiB buffer,
// the last one at the end. This prints the final line of that sequence.
void reportCompressed(size_t inputBytes, uint64_t positions, bool binInput)
{
    const uint64_t reported = positions / 100000 * 100000;
    if (reported == 0) return;
    if (binInput) std::cout << "Processed " << reported * 40 << " bytes and " << reported << " positions.\n";
    else if (reported == positions) std::cout << "Processed " << inputBytes << " bytes and " << reported << " positions.\n";
}
void reportDecompressed(size_t outputBytes, uint64_t positions)
{
    if (outputBytes > 0) std::cout << "Processed " << outputBytes << " bytes and " << positions << " positions.\n";
}

void runDriver(driver_fn fn, const char* verb, bool compresses, const std::string& inputPath, const std::string& outputPath,
               bool append)
{
    std::cout << verb << " " << inputPath << " to " << outputPath << '\n';
    std::ifstream in(inputPath, std::ios_base::binary | std::ios_base::ate);
    const size_t n = in ? static_cast<size_t>(in.tellg()) : 0;
    HostBuffer src(n);
    if (!src.p) fail(NNP_ERR_NOMEM);
    in.seekg(0);
    in.read(src.p, static_cast<std::streamsize>(n));

    size_t need = 0;
    int rc = fn(src.p, n, nullptr, 0, &need);  // size query (a count pass for the directions whose output is variable)
    if (rc != NNP_OK && !isReferenceError(rc)) fail(rc);
    HostBuffer dst(need);
    if (!dst.p) fail(NNP_ERR_NOMEM);
    size_t produced = 0;
    rc = fn(src.p, n, dst.p, need, &produced);
    if (rc != NNP_OK && !isReferenceError(rc)) fail(rc);
    {
        std::ofstream out(outputPath, std::ios_base::binary | (append ? std::ios_base::app : std::ios_base::trunc));
        out.write(dst.p, static_cast<std::streamsize>(produced));
        if (!out) fail(NNP_ERR_BAD_ARG);
    }
    if (isReferenceError(rc)) throw std::runtime_error(nnp_strerror(rc));  // printed by main, exit code 0
    if (compresses) reportCompressed(n, nnp_last_positions(), endsWith(inputPath, binExtension));
    else reportDecompressed(produced, nnp_last_positions());
}

// The two headline directions go file to file through the slab-wise drivers (nnp_*_file: reader thread,
// H2D, kernels, D2H and writer overlap; neither host nor device memory has to hold the whole file) as
// soon as the larger of input and expected output exceeds NNP_STREAM_THRESHOLD bytes (default 256 MiB;
// a .binpack is taken to expand 24 times). NNP_SLAB_BYTES sets the slab size (0 = the library's default).
typedef int (*file_fn)(const char*, const char*, int, size_t, uint64_t*);

bool runFileDriver(file_fn fn, const char* verb, bool compresses, const std::string& inputPath, const std::string& outputPath,
                   bool append)
{
    std::ifstream in(inputPath, std::ios_base::binary | std::ios_base::ate);
    const unsigned long long size = in ? static_cast<unsigned long long>(in.tellg()) : 0ull;
    const char* t = std::getenv("NNP_STREAM_THRESHOLD");
    const unsigned long long threshold = t ? std::strtoull(t, nullptr, 10) : (256ull << 20);
    if ((compresses ? size : size * 24) <= threshold) return false;
    const char* sl = std::getenv("NNP_SLAB_BYTES");
    std::cout << verb << " " << inputPath << " to " << outputPath << '\n';
    uint64_t positions = 0;
    const int rc = fn(inputPath.c_str(), outputPath.c_str(), append ? 1 : 0, sl ? std::strtoull(sl, nullptr, 10) : 0, &positions);
    if (isReferenceError(rc)) throw std::runtime_error(nnp_strerror(rc));
    if (rc != NNP_OK) fail(rc);
    if (compresses) reportCompressed(size, positions, true);
    else reportDecompressed(positions * 40, positions);
    return true;
}

void convert(const std::string& inputPath, std::string outputPath, bool append)
{
    if (!fileExists(inputPath)) {
        std::cerr << "Input file doesn't exist.\n";
        return;
    }
    if (endsWith(inputPath, binExtension) && endsWith(outputPath, plainExtension)) {
        runDriver(nnp_bin_to_plain, "Converting", false, inputPath, outputPath, append);
    } else if (endsWith(inputPath, plainExtension) && endsWith(outputPath, binExtension)) {
        runDriver(nnp_plain_to_bin, "Compressing", true, inputPath, outputPath, append);
    } else if (endsWith(inputPath, plainExtension) || endsWith(inputPath, binExtension)) {
        if (!endsWith(outputPath, binpackExtension)) outputPath += binpackExtension;
        if (endsWith(inputPath, binExtension) && runFileDriver(nnp_bin_to_binpack_file, "Compressing", true, inputPath, outputPath, append))
            return;
        runDriver(endsWith(inputPath, binExtension) ? nnp_bin_to_binpack : nnp_plain_to_binpack, "Compressing", true, inputPath,
                  outputPath, append);
    } else if (endsWith(inputPath, binpackExtension)) {
        if (endsWith(outputPath, binExtension)) {
            if (runFileDriver(nnp_binpack_to_bin_file, "Decompressing", false, inputPath, outputPath, append)) return;
            runDriver(nnp_binpack_to_bin, "Decompressing", false, inputPath, outputPath, append);
        } else if (endsWith(outputPath, plainExtension)) {
            runDriver(nnp_binpack_to_plain, "Decompressing", false, inputPath, outputPath, append);
        } else {
            std::cerr << "Unrecognized file format. Only " << binExtension << " and " << plainExtension
                      << " are supported for decompression.";
        }
    } else {
        std::cerr << "Unsupported extension.";
    }
}

void help()
{
    std::cout << "Usage:\n";
    std::cout << "    nnue_data_compression [-h] [-a] input_path output_path\n";
    std::cout << "\n";
    std::cout << "-h, --help                show help\n";
    std::cout << "-a, --append              append to the output file instead of truncating it\n";
    std::cout << "\n";
    std::cout << "input_path                the path to the file to process\n";
    std::cout << "output_path               the path to the file to create/append to\n";
    std::cout << "\n";
    std::cout << "Behaviour depends on file extensions. If the input\n";
    std::cout << "file has extension either " << binExtension << " or " << plainExtension << "\n";
    std::cout << "it will be compressed. The output file has then an implied\n";
    std::cout << "extension of " << binpackExtension << " and it doesn't have to be specified.\n";
    std::cout << "If the input file's extension is " << binpackExtension << " then it will be decompressed\n";
    std::cout << "to either a " << binExtension << " or " << plainExtension << " file, depending on the extension.\n";
    std::cout << "\n";
    std::cout << "Example usage:\n";
    std::cout << "1. convert from plain to binpack in append mode:\n";
    std::cout << "    nnue_data_compression -a data.plain data\n";
    std::cout << "2. convert from binpack to plain in truncate/replace mode:\n";
    std::cout << "    nnue_data_compression data.binpack data.plain\n";
}

}  // namespace

int main(int argc, char** argv)
{
    std::set<std::string> flags;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        if (*argv[i] == '-') flags.emplace(argv[i] + 1);
        else pos.emplace_back(argv[i]);
    }
    if (pos.empty() || flags.count("help") == 1 || flags.count("h") == 1) {
        help();
        return 0;
    }
    if (pos.size() != 2) {
        std::cerr << "Invalid arguments.\n";
        help();
        return 1;
    }
    const bool append = flags.count("a") == 1 || flags.count("append") == 1;
    // NNP_DEVICES = "all" or a count: every slab-wise conversion is spread over that many GPUs (one loader,
    // compute and drainer thread per GPU, csrc/files.cu); otherwise one GPU, NNP_DEVICE (default 0)
    const char* devs = std::getenv("NNP_DEVICES");
    const char* dev = std::getenv("NNP_DEVICE");
    int rc;
    if (devs && *devs) {
        rc = nnp_init_all(std::string(devs) == "all" ? 0 : std::atoi(devs));
        if (rc > 0) rc = NNP_OK;
    } else {
        rc = nnp_init(dev ? std::atoi(dev) : 0);
    }
    if (rc != NNP_OK) {
        std::cerr << nnp_strerror(rc) << " " << nnp_last_cuda_error() << "\n";
        return 2;
    }
    try {
        convert(pos[0], pos[1], append);
    } catch (std::runtime_error& e) {
        std::cerr << e.what() << "\n";
        std::cerr << "Exiting...\n";
    }
    nnp_shutdown();
    return 0;
}
