// Legacy-ASCII VTK output and input, and the progress line (reference:
// include/vtk.cuh; format: http://www.vtk.org/wp-content/uploads/2015/04/
// file-formats.pdf).
//
// Vtk_output writes one file per call of write_positions,
// <output_path><base_name>_<frame>.vtk, and appends the further sections
// (links, fields, polarities, properties) to that file. Vtk_input reads the
// sections back by keyword. Host-only; not part of the timed step.
#pragma once

#include <assert.h>
#include <stdint.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <vector>

#include "links.cuh"
#include "polarity.cuh"
#include "utils.cuh"


template<typename Pt, template<typename> class Solver>
class Solution;

template<typename Prop>
struct Property;


class Vtk_output {
public:
    // Files are stored as output_path/base_name_#.vtk
    Vtk_output(std::string base_name, std::string output_path = "output/",
        bool verbose = true);
    ~Vtk_output(void);
    // Write x, y, and z component of Pt; has to be written first. Points with
    // input_mask[i] == false are left out of this and all following sections.
    template<typename Pt, template<typename> class Solver>
    void write_positions(Solution<Pt, Solver>& points, bool* input_mask = NULL);
    // Write links, see links.cuh; if written has to be second
    void write_links(Links& links);
    // Write further components of Pt
    template<typename Pt, template<typename> class Solver>
    void write_field(Solution<Pt, Solver>& points, const char* data_name = "w",
        float Pt::*field = &Pt::w);
    // Write a polarity of Pt as unit normals, see polarity.cuh; the default
    // theta = phi = 0 is written as {0, 0, 0}.
    template<typename Pt, float Pt::*theta = &Pt::theta,
        float Pt::*phi = &Pt::phi, template<typename> class Solver>
    void write_polarity(
        Solution<Pt, Solver>& points, const char* data_name = "polarity");
    // Write not integrated property, see property.cuh
    template<typename Prop>
    void write_property(Property<Prop>& property);

private:
    int n_points = 0;
    int n_to_write = 0;
    bool* mask = NULL;
    int time_step = 0;
    std::string base_name;
    std::string output_dir;
    std::string current_path;
    bool verbose;
    bool point_data_started = false;
    time_t t_0;

    bool skipped(int i) const { return mask != NULL && !mask[i]; }

    // Re-open the current frame for appending; the first POINT_DATA section
    // writer also emits the section header.
    std::ofstream append_point_data()
    {
        std::ofstream file(current_path, std::ios_base::app);
        assert(file.is_open());
        if (!point_data_started) {
            file << "\nPOINT_DATA " << n_to_write << "\n";
            point_data_started = true;
        }
        return file;
    }
};

inline Vtk_output::Vtk_output(
    std::string base_name, std::string output_path, bool verbose)
    : base_name{base_name}, output_dir{output_path}, verbose{verbose}
{
    if (output_dir.empty() || output_dir.back() != '/') {
        output_dir.append("/");
        std::cout << output_dir << std::endl;
    }
    mkdir(output_dir.c_str(), 0755);
    time(&t_0);
}

inline Vtk_output::~Vtk_output()
{
    if (!verbose) return;

    const auto duration = time(NULL) - t_0;
    std::cout << "Integrating " << base_name << ", ";
    if (duration < 60)
        std::cout << duration << " seconds";
    else if (duration < 60 * 60)
        std::cout << duration / 60 << "m " << duration % 60 << "s";
    else
        std::cout << duration / (60 * 60) << "h " << duration % (60 * 60)
                  << "m";
    std::cout << " taken (" << n_points << " points).        \n";
}

template<typename Pt, template<typename> class Solver>
void Vtk_output::write_positions(Solution<Pt, Solver>& points, bool* input_mask)
{
    n_points = *points.h_n;
    mask = input_mask;
    n_to_write = 0;
    for (int i = 0; i < n_points; i++) n_to_write += skipped(i) ? 0 : 1;

    current_path =
        output_dir + base_name + "_" + std::to_string(time_step) + ".vtk";
    std::ofstream file(current_path);
    assert(file.is_open());

    file << "# vtk DataFile Version 3.0\n"
         << base_name << "\n"
         << "ASCII\n"
         << "DATASET POLYDATA\n"
         << "\nPOINTS " << n_to_write << " float\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        const Pt& X = points.h_X[i];
        file << X.x << " " << X.y << " " << X.z << "\n";
    }

    file << "\nVERTICES " << n_to_write << " " << 2 * n_to_write << "\n";
    for (int i = 0; i < n_to_write; i++) file << "1 " << i << "\n";

    point_data_started = false;
    time_step += 1;
    if (!verbose) return;

    std::cout << "Integrating " << base_name << ", " << time_step
              << " steps done (" << n_points << " points)        \r";
    std::cout.flush();
}

inline void Vtk_output::write_links(Links& links)
{
    std::ofstream file(current_path, std::ios_base::app);
    assert(file.is_open());

    const int n_links = *links.h_n;
    file << "\nLINES " << n_links << " " << 3 * n_links << "\n";
    for (int i = 0; i < n_links; i++)
        file << "2 " << links.h_link[i].a << " " << links.h_link[i].b << "\n";
}

template<typename Pt, template<typename> class Solver>
void Vtk_output::write_field(
    Solution<Pt, Solver>& points, const char* data_name, float Pt::*field)
{
    std::ofstream file = append_point_data();
    file << "SCALARS " << data_name << " float\n"
         << "LOOKUP_TABLE default\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        file << points.h_X[i].*field << "\n";
    }
}

template<typename Pt, float Pt::*theta, float Pt::*phi,
    template<typename> class Solver>
void Vtk_output::write_polarity(
    Solution<Pt, Solver>& points, const char* data_name)
{
    std::ofstream file = append_point_data();
    file << "NORMALS " << data_name << " float\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        const Pt& X = points.h_X[i];
        float3 n = pol_to_float3<Pt, theta, phi>(X);
        if (X.*theta == 0 && X.*phi == 0) n.z = 0;  // "no polarity"
        file << n.x << " " << n.y << " " << n.z << "\n";
    }
}

template<typename Prop>
void Vtk_output::write_property(Property<Prop>& property)
{
    std::ofstream file = append_point_data();
    // float properties are written as floats, everything else as int
    const std::string type_name =
        std::string(typeid(Prop).name()) == "f" ? "float" : "int";

    assert(n_points <= property.n_max);
    file << "SCALARS " << property.name << " " << type_name << "\n"
         << "LOOKUP_TABLE default\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        file << property.h_prop[i] << "\n";
    }
}


// Reads frames back by section keyword. Besides the ASCII frames of Vtk_output
// it understands the BINARY frames of Vtk_async_output (big-endian payloads):
// for those the sections are indexed once, by walking the file from payload to
// payload, so that binary data is never mistaken for a keyword line.
class Vtk_input {
public:
    Vtk_input(std::string file_name);
    // Stream position just behind the line that starts with the two keywords
    std::streampos find_entry(std::string, std::string);
    template<typename Pt, template<typename> class Solver>
    void read_positions(Solution<Pt, Solver>& points);
    // Read polarity of Pt, see polarity.cuh
    template<typename Pt, template<typename> class Solver>
    void read_polarity(Solution<Pt, Solver>& points);
    // Read further field of Pt
    template<typename Pt, template<typename> class Solver>
    void read_field(Solution<Pt, Solver>& points, const char* data_name = "w",
        float Pt::*field = &Pt::w);
    // Read property, see property.cuh
    template<typename Prop>
    void read_property(Property<Prop>& property, std::string prop_name);
    int n_points;

private:
    struct Section {
        std::string keyword, name;
        std::streampos data;  // first byte behind the keyword line(s)
    };
    std::string file_name;
    bool binary = false;
    std::vector<Section> sections;  // binary frames only

    void index_binary_sections();

    // Hands the next `rows` records of `width` values of type T to
    // take(i, values): one text line per record in ASCII frames (parsed like
    // the reference does, with operator>> of T), 32-bit big-endian words
    // otherwise.
    template<typename T, typename Take>
    void for_each_record(std::string keyword1, std::string keyword2,
        int header_lines, int rows, int width, Take take)
    {
        static_assert(sizeof(T) == 4, "legacy VTK float / int payloads");
        std::ifstream file(file_name, std::ios::binary);
        assert(file.is_open());
        file.seekg(find_entry(keyword1, keyword2));
        std::vector<T> values(width);
        if (binary) {
            std::vector<uint32_t> words(width);
            for (int i = 0; i < rows; i++) {
                file.read(reinterpret_cast<char*>(words.data()),
                    width * sizeof(uint32_t));
                for (int k = 0; k < width; k++) {
                    const uint32_t bits = __builtin_bswap32(words[k]);
                    memcpy(&values[k], &bits, sizeof(T));
                }
                take(i, values);
            }
            return;
        }
        std::string line;
        for (int i = 0; i < header_lines; i++) getline(file, line);
        for (int i = 0; i < rows; i++) {
            getline(file, line);
            std::istringstream text(line);
            for (int k = 0; k < width; k++) text >> values[k];
            take(i, values);
        }
    }
};

inline Vtk_input::Vtk_input(std::string file_name) : file_name{file_name}
{
    std::ifstream file(file_name, std::ios::binary);
    assert(file.is_open());

    // line 3 names the encoding, "POINTS <n> float" is among the first six
    n_points = 0;
    std::string line;
    for (int i = 0; i < 6; i++) {
        getline(file, line);
        if (i == 2) binary = line.compare(0, 6, "BINARY") == 0;
        const auto items = split(line);
        if (items.size() > 1 && items[0] == "POINTS") {
            n_points = stoi(items[1]);
            break;
        }
    }
    if (binary) index_binary_sections();
}

inline void Vtk_input::index_binary_sections()
{
    std::ifstream file(file_name, std::ios::binary);
    std::string line;
    for (int i = 0; i < 4; i++) getline(file, line);  // header
    while (getline(file, line)) {
        const auto items = split(line);
        if (items.size() < 2) continue;  // blank separator lines
        size_t payload = 0;
        Section section{items[0], items[1], 0};
        if (items[0] == "POINTS" || items[0] == "NORMALS") {
            payload = size_t(n_points) * 3 * sizeof(float);
        } else if (items[0] == "VERTICES") {
            payload = size_t(stoi(items[1])) * 2 * sizeof(uint32_t);
        } else if (items[0] == "LINES" && items.size() > 2) {
            payload = size_t(stoi(items[2])) * sizeof(uint32_t);
        } else if (items[0] == "SCALARS") {
            getline(file, line);  // LOOKUP_TABLE default
            payload = size_t(n_points) * sizeof(float);
        } else if (items[0] == "POINT_DATA") {
            continue;
        } else {
            break;  // unknown section: stop indexing rather than guess a size
        }
        section.data = file.tellg();
        sections.push_back(section);
        file.seekg(section.data + std::streamoff(payload));
    }
}

inline std::streampos Vtk_input::find_entry(
    std::string keyword1, std::string keyword2)
{
    if (binary) {
        for (const auto& section : sections)
            if (section.keyword == keyword1 && section.name == keyword2)
                return section.data;
    } else {
        std::ifstream file(file_name);
        assert(file.is_open());
        std::string line;
        for (int i = 0; i < 4; i++) getline(file, line);  // header
        while (getline(file, line)) {
            const auto items = split(line);
            if (items.size() > 1 && items[0] == keyword1 && items[1] == keyword2)
                return file.tellg();
        }
    }
    std::cout << "Vtk_input: no entry \"" << keyword1 << " " << keyword2
              << "\" in " << file_name << std::endl;
    assert(false);
    return std::streampos(0);
}

template<typename Pt, template<typename> class Solver>
void Vtk_input::read_positions(Solution<Pt, Solver>& points)
{
    for_each_record<float>("POINTS", std::to_string(n_points), 0, n_points, 3,
        [&](int i, const std::vector<float>& v) {
            points.h_X[i].x = v[0];
            points.h_X[i].y = v[1];
            points.h_X[i].z = v[2];
        });
}

template<typename Pt, template<typename> class Solver>
void Vtk_input::read_polarity(Solution<Pt, Solver>& points)
{
    for_each_record<float>("NORMALS", "polarity", 0, n_points, 3,
        [&](int i, const std::vector<float>& v) {
            const float x = v[0], y = v[1], z = v[2];
            const auto length = sqrt(pow(x, 2) + pow(y, 2) + pow(z, 2));
            if (length == 0) {  // written for theta = phi = 0
                points.h_X[i].phi = 0.0f;
                points.h_X[i].theta = 0.0f;
            } else {  // the normals are unit vectors
                points.h_X[i].phi = atan2(y, x);
                points.h_X[i].theta = acos(z);
            }
        });
}

template<typename Pt, template<typename> class Solver>
void Vtk_input::read_field(
    Solution<Pt, Solver>& points, const char* data_name, float Pt::*field)
{
    for_each_record<float>("SCALARS", data_name, 1, n_points, 1,  // LOOKUP_TABLE
        [&](int i, const std::vector<float>& v) { points.h_X[i].*field = v[0]; });
}

template<typename Prop>
void Vtk_input::read_property(Property<Prop>& property, std::string prop_name)
{
    assert(n_points <= property.n_max);
    // float properties are stored as floats, everything else as int
    using Stored = typename std::conditional<std::is_same<Prop, float>::value,
        float, int>::type;
    for_each_record<Stored>("SCALARS", prop_name, 1, n_points, 1,
        [&](int i, const std::vector<Stored>& v) {
            property.h_prop[i] = static_cast<Prop>(v[0]);
        });
}


// Extension: Vtk_async_output<Pt>, snapshots on the solver's stream and writes
// frames (ASCII or binary) from a background thread.
#include "b200/vtk_async.cuh"
