"""Interactive client for the REST text-generation server (parity: tools/text_generation_cli.py).

usage: python tools/text_generation_cli.py host:port"""
import json
import sys
import urllib.request


def query(url, prompt, tokens_to_generate):
    req = urllib.request.Request(url, data=json.dumps({"prompts": [prompt], "tokens_to_generate": tokens_to_generate})
                                 .encode(), headers={"Content-Type": "application/json; charset=UTF-8"}, method="PUT")
    with urllib.request.urlopen(req) as resp:
        return json.loads(resp.read())


if __name__ == "__main__":
    url = "http://" + sys.argv[1] + "/api"
    while True:
        sentence = input("Enter prompt: ")
        tokens_to_generate = int(input("Enter number of tokens to generate: "))
        print("Megatron Response: ")
        print(query(url, sentence, tokens_to_generate)["text"][0])
