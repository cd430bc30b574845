/* Test infrastructure only (oracle/): stand-in for <gnuradio/msg_queue.h> + <gnuradio/message.h>. */
#ifndef ORACLE_SHIM_GNURADIO_MSG_QUEUE_H
#define ORACLE_SHIM_GNURADIO_MSG_QUEUE_H
#include <memory>
#include <string>
#include <vector>
namespace gr {
class message {
public:
    typedef std::shared_ptr<message> sptr;
    static sptr make_from_string(const std::string& s) {
        sptr m = std::make_shared<message>(); m->text = s; return m;
    }
    std::string to_string() const { return text; }
    std::string text;
};
class msg_queue {
public:
    typedef std::shared_ptr<msg_queue> sptr;
    void handle(message::sptr m) { msgs.push_back(m->to_string()); }
    std::vector<std::string> msgs;
};
}  // namespace gr
#endif
