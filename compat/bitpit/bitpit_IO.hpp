// bitpit_IO.hpp -- minimal stand-in for the parts of bitpit's IO module minimmerflow uses
// (see README.md): the global XML configuration tree `config::root` and the `log::cout()` logger.
#ifndef MMF_COMPAT_BITPIT_IO_HPP
#define MMF_COMPAT_BITPIT_IO_HPP

#include "bitpit_common.hpp"

namespace bitpit {

// ---- configuration tree -------------------------------------------------------------------------
// An XML element with child elements is a section, an element with text only is an option.
class Config {
public:
    typedef std::multimap<std::string, std::unique_ptr<Config>> Sections;
    typedef std::map<std::string, std::string> Options;

    bool hasOption(const std::string &key) const { return m_options.count(key) > 0; }
    bool hasSection(const std::string &key) const { return m_sections.count(key) > 0; }
    const Sections &getSections() const { return m_sections; }
    const Options &getOptions() const { return m_options; }

    Config &getSection(const std::string &key)
    {
        auto it = m_sections.find(key);
        if (it == m_sections.end()) throw std::runtime_error("The section named \"" + key + "\" does not exist");
        return *it->second;
    }
    const Config &getSection(const std::string &key) const { return const_cast<Config *>(this)->getSection(key); }
    Config &operator[](const std::string &key) { return getSection(key); }
    const Config &operator[](const std::string &key) const { return getSection(key); }

    const std::string &get(const std::string &key) const
    {
        auto it = m_options.find(key);
        if (it == m_options.end()) throw std::runtime_error("The option named \"" + key + "\" does not exist");
        return it->second;
    }
    template <typename T>
    T get(const std::string &key) const
    {
        return convert<T>(get(key));
    }
    template <typename T>
    T get(const std::string &key, const T &fallback) const
    {
        return hasOption(key) ? convert<T>(get(key)) : fallback;
    }

    Config &addSection(const std::string &key) { return *m_sections.emplace(key, std::unique_ptr<Config>(new Config()))->second; }
    void set(const std::string &key, const std::string &value) { m_options[key] = value; }
    void clear() { m_options.clear(); m_sections.clear(); }

private:
    Options m_options;
    Sections m_sections;

    template <typename T>
    static T convert(const std::string &text)
    {
        std::istringstream s(text);
        T value;
        s >> value;
        if (s.fail()) throw std::runtime_error("cannot convert configuration value \"" + text + "\"");
        return value;
    }
};

template <>
inline std::string Config::convert<std::string>(const std::string &text)
{
    const std::size_t a = text.find_first_not_of(" \t\r\n"), b = text.find_last_not_of(" \t\r\n");
    return (a == std::string::npos) ? std::string() : text.substr(a, b - a + 1);
}

namespace config {

namespace detail {

// a deliberately small XML reader: declaration, comments, nested elements, text; attributes are skipped
class XmlReader {
public:
    explicit XmlReader(const std::string &text) : m_text(text) {}

    void parseDocument(Config *root, std::string *rootName)
    {
        skipMisc();
        std::string name;
        bool selfClosed;
        if (!openTag(&name, &selfClosed)) throw std::runtime_error("configuration file: no root element");
        *rootName = name;
        if (!selfClosed) parseContent(name, root, nullptr);
    }

private:
    const std::string &m_text;
    std::size_t m_pos = 0;

    void skipMisc()
    {
        for (;;) {
            while (m_pos < m_text.size() && std::isspace((unsigned char) m_text[m_pos])) ++m_pos;
            if (m_text.compare(m_pos, 4, "<!--") == 0) {
                const std::size_t e = m_text.find("-->", m_pos);
                if (e == std::string::npos) throw std::runtime_error("configuration file: unterminated comment");
                m_pos = e + 3;
            } else if (m_text.compare(m_pos, 2, "<?") == 0) {
                const std::size_t e = m_text.find("?>", m_pos);
                if (e == std::string::npos) throw std::runtime_error("configuration file: unterminated declaration");
                m_pos = e + 2;
            } else {
                return;
            }
        }
    }

    bool openTag(std::string *name, bool *selfClosed)
    {
        if (m_pos >= m_text.size() || m_text[m_pos] != '<' || m_text.compare(m_pos, 2, "</") == 0) return false;
        const std::size_t e = m_text.find('>', m_pos);
        if (e == std::string::npos) throw std::runtime_error("configuration file: unterminated tag");
        std::string inside = m_text.substr(m_pos + 1, e - m_pos - 1);
        *selfClosed = !inside.empty() && inside.back() == '/';
        if (*selfClosed) inside.pop_back();
        const std::size_t sp = inside.find_first_of(" \t\r\n");
        *name = inside.substr(0, sp);
        m_pos = e + 1;
        return true;
    }

    // content of element `name` up to its closing tag; fills `section` when children are found,
    // otherwise returns the text through `text`
    void parseContent(const std::string &name, Config *section, std::string *text)
    {
        std::string collected;
        for (;;) {
            const std::size_t lt = m_text.find('<', m_pos);
            if (lt == std::string::npos) throw std::runtime_error("configuration file: element <" + name + "> is not closed");
            collected += m_text.substr(m_pos, lt - m_pos);
            m_pos = lt;
            if (m_text.compare(m_pos, 4, "<!--") == 0 || m_text.compare(m_pos, 2, "<?") == 0) {
                skipMisc();
                continue;
            }
            if (m_text.compare(m_pos, 2, "</") == 0) {
                const std::size_t e = m_text.find('>', m_pos);
                m_pos = (e == std::string::npos) ? m_text.size() : e + 1;
                if (text) *text = collected;
                return;
            }
            std::string child;
            bool selfClosed;
            openTag(&child, &selfClosed);
            if (selfClosed) {
                section->set(child, "");
                continue;
            }
            // look ahead: does the child hold elements (section) or text (option)?
            Config probe;
            std::string childText;
            const std::size_t save = m_pos;
            XmlReader sub(m_text);
            sub.m_pos = save;
            sub.parseContent(child, &probe, &childText);
            m_pos = sub.m_pos;
            if (probe.getSections().empty() && probe.getOptions().empty()) {
                section->set(child, childText);
            } else {
                Config &target = section->addSection(child);
                XmlReader again(m_text);
                again.m_pos = save;
                again.parseContent(child, &target, nullptr);
            }
        }
    }
};

} // namespace detail

class GlobalConfigParser : public Config {
public:
    void reset(const std::string &rootName, int version) { clear(); m_rootName = rootName; m_version = version; }
    void read(const std::string &fileName)
    {
        std::ifstream in(fileName);
        if (!in) throw std::runtime_error("Unable to read the configuration file \"" + fileName + "\"");
        std::stringstream buffer;
        buffer << in.rdbuf();
        const std::string text = buffer.str();
        std::string rootName;
        detail::XmlReader(text).parseDocument(this, &rootName);
        if (!m_rootName.empty() && rootName != m_rootName) {
            throw std::runtime_error("The name of the root element of \"" + fileName + "\" is not \"" + m_rootName + "\"");
        }
    }

private:
    std::string m_rootName;
    int m_version = 0;
};

inline GlobalConfigParser &rootInstance()
{
    static GlobalConfigParser instance;
    return instance;
}
static GlobalConfigParser &root = rootInstance();

inline void reset(const std::string &rootName, int version) { root.reset(rootName, version); }
inline void read(const std::string &fileName) { root.read(fileName); }

} // namespace config

// ---- logger -------------------------------------------------------------------------------------
namespace log {

enum Mode { SEPARATE, COMBINED };
enum Visibility { MASTER, GLOBAL };

// everything written to the logger goes to the console and to "<directory>/<name>.log"
class TeeBuffer : public std::streambuf {
public:
    void open(const std::string &path) { m_file.open(path, std::ios::out | std::ios::trunc); }

protected:
    int overflow(int ch) override
    {
        if (ch != EOF) {
            std::cout.put((char) ch);
            if (m_file.is_open()) m_file.put((char) ch);
        }
        return ch;
    }
    int sync() override
    {
        std::cout.flush();
        if (m_file.is_open()) m_file.flush();
        return 0;
    }

private:
    std::ofstream m_file;
};

class Logger : public std::ostream {
public:
    Logger() : std::ostream(&m_buffer) {}
    void setVisibility(Visibility) {}
    void setDefaultVisibility(Visibility) {}
    TeeBuffer &buffer() { return m_buffer; }

private:
    TeeBuffer m_buffer;
};

class LoggerManager {
public:
    void initialize(Mode, const std::string &name, bool reset, const std::string &directory, int nProcessors, int rank)
    {
        BITPIT_UNUSED(reset);
        BITPIT_UNUSED(nProcessors);
        BITPIT_UNUSED(rank);
        m_logger.buffer().open(directory + "/" + name + ".log");
    }
    Logger &cout() { return m_logger; }

private:
    Logger m_logger;
};

inline LoggerManager &manager()
{
    static LoggerManager instance;
    return instance;
}

inline Logger &cout() { return manager().cout(); }

} // namespace log

} // namespace bitpit

#endif
