"""Fuzz inputs for the two byte-level parsers of the path: the ELB hostname tokeniser and the listen-ports JSON."""
import random


def fuzz_hostnames(seed: int, n: int = 2000):
    """lbIngress hostnames built from a small adversarial alphabet around the ELB patterns."""
    rng = random.Random(seed)
    alpha = "abcXYZ019_-." + chr(10) + "~ " + chr(0xE9)
    hosts = []
    for _ in range(n):
        kind = rng.random()
        core = "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 14)))
        if kind < 0.3:
            h = core + ".elb.us-east-1.amazonaws.com"
        elif kind < 0.6:
            h = core + ".us-east-1.elb.amazonaws.com"
        elif kind < 0.7:
            h = "internal-" + core + ".eu.elb.amazonaws.com"
        elif kind < 0.8:
            h = core + ".elb." + "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 5))) + ".amazonaws.com"
        elif kind < 0.9:
            h = core + ".amazonaws.com"
        else:
            h = core
        hosts.append(h)
    return [dict(kind="service", ns="d", name=f"s{i}", annotations={}, lb_ingress=[h]) for i, h in enumerate(hosts)]


_CTRL = chr(1)
LISTEN_FUZZ = [
    '[{"HTTP": 80}, {"HTTPS": 443}]', '[{"https":443,"HTTP":8080}]', '[{"HTTP": 80', '[{"HTTP": 80}] ', '\t[\n{"HTTPS":\r443}\n]', '[{"HTTP":1e3}]',
    '[{"HTTP":-0}]', '[{"HTTP":00}]', '[{"x":{"a":[1,2,{"b":"\\u00e9\\n"}]},"HTTPS":8443}]', '[{"HTTP":80,"HTTP":null}]', '[{"HTTP":null,"HTTP":81}]',
    '[{"h\\u0074tp":8}]', '[{"HTTP\\u017f":9}]', '[{"HTTPS":9223372036854775807}]', '[{"HTTPS":9223372036854775808}]', '[{"HTTPS":-9223372036854775808}]',
    '[{"HTTP":2147483648}]', '[null,null,{"HTTP":1},null]', '[{}]', '[[]]', '[1]', '["HTTP"]', 'nul', 'null ', ' null', '[{"HTTP":80}],', '[{"HTTP":80},]',
    '[{"HTTP":80,}]', '[{,"HTTP":80}]', '[{"HTTP" 80}]', '[{"HTTP":80 "HTTPS":1}]', '[{"a":"\\ud83d\\ude00","HTTP":7}]', '[{"a":"\\ud83d","HTTP":7}]',
    '[{"a":"' + _CTRL + '","HTTP":7}]', '[{"a":tru,"HTTP":7}]', '[{"a":true,"b":false,"c":null,"HTTP":7}]', '[{"a":-,"HTTP":7}]', '[{"a":1.5e+3,"HTTP":7}]',
    '[{"a":1.e3,"HTTP":7}]', '[{"a":.5,"HTTP":7}]', '[' * 50 + ']' * 50, '[{"a":' + '[' * 300 + ']' * 300 + ',"HTTPS":5}]', '[{"HTTP":80}]x', '',
    '[{"http' + chr(0x17F) + '": 4443}]', '[{"HTTP' + chr(0x212A) + '": 1}]',
]


def listen_objects():
    return [dict(kind="ingress", ns="d", name=f"i{k}", ingress_class="alb", annotations={"alb.ingress.kubernetes.io/listen-ports": v}, ports=[80])
            for k, v in enumerate(LISTEN_FUZZ)]
